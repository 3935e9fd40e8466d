"""Times md_attention_bf16 alone at the bench shape (CUDA events, warm, B sequences): python tools/att_bench.py 64 [L]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musediffusion_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 2096
NH = 12
qkv = (torch.randn(B * L, 3 * NH * 64, device="cuda") * 0.7).to(torch.bfloat16)
out = torch.empty(B * L, NH * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, B, L, NH, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    ops.attention(qkv, B, L, NH, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("attention B=%d L=%d: %.3f ms  %.1f TFLOP/s  (MD_ATT_POLY=%s MD_ATT_THR=%s)"
      % (B, L, ms, 4.0 * B * NH * L * L * 64 / ms / 1e9, os.environ.get("MD_ATT_POLY", "default"), os.environ.get("MD_ATT_THR", "default")))
