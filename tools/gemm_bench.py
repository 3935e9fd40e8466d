"""Times md_linear_bf16 at the denoiser shapes (CUDA events): python tools/gemm_bench.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musediffusion_b200 import _lib, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M = B * 2096
for name, N, K, epi in [("qkv", 2304, 768, _lib.EPI_BIAS), ("out", 768, 768, _lib.EPI_BIAS), ("ffn1+gelu", 3072, 768, _lib.EPI_BIAS_GELU), ("ffn2", 768, 3072, _lib.EPI_BIAS)]:
    A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.linear(A, W, b, epi, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.linear(A, W, b, epi, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-10s M=%d N=%d K=%d: %.3f ms  %.0f TFLOP/s  (MD_GEMM_DEBUG_SKIP=%s)" % (name, M, N, K, ms, 2.0 * M * N * K / ms / 1e9, os.environ.get("MD_GEMM_DEBUG_SKIP", "0")))
    del A, W, out
