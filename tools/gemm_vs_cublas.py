"""md_linear_bf16 against cuBLAS (torch.matmul / F.linear, bf16) on the four encoder GEMM shapes of the bench, each in a loop
for a few seconds so that both run at the board's power limit: sustained TFLOP/s, clock, power.  The cuBLAS side has no
bias / GELU epilogue (a lower bound on its time for the same work).  python tools/gemm_vs_cublas.py [--batch 256]"""
import argparse, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from musediffusion_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--seconds", type=float, default=2.5)
a = ap.parse_args()
dev = torch.device("cuda:0")
M = a.batch * 2096
bf = torch.bfloat16


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop = [], False

    def run(self):
        while not self.stop:
            out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=power.draw,clocks.sm", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True).stdout.strip().split(",")
            if len(out) >= 2:
                self.rows.append((float(out[0]), float(out[1])))
            time.sleep(0.1)


def measure(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = Sampler()
    s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, tic = 0, time.perf_counter()
    e0.record()
    while time.perf_counter() - tic < a.seconds:
        for _ in range(8):
            fn()
        n += 8
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    s.stop = True
    s.join()
    rows = sorted(s.rows[len(s.rows) // 3:])
    w = rows[len(rows) // 2][0]
    mhz = sorted(r[1] for r in s.rows[len(s.rows) // 3:])[len(rows) // 2]
    return e0.elapsed_time(e1) / n, w, mhz


lines = []
for name, N, K, epi in (("QKV 768 -> 2304", 2304, 768, _lib.EPI_BIAS), ("out-proj 768 -> 768", 768, 768, _lib.EPI_BIAS),
                        ("FFN1 768 -> 3072 (+GELU in ours)", 3072, 768, _lib.EPI_BIAS_GELU), ("FFN2 3072 -> 768", 768, 3072, _lib.EPI_BIAS)):
    x = (torch.randn(M, K, device=dev) * 0.5).to(bf)
    W = (torch.randn(N, K, device=dev) * 0.03).to(bf)
    b = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=bf)
    fl = 2.0 * M * N * K
    ms_o, w_o, f_o = measure(lambda: ops.linear(x, W, b, epilogue=epi, out=out))
    ms_c, w_c, f_c = measure(lambda: torch.matmul(x, W.t(), out=out))
    line = ("%-34s ours %6.3f ms %6.0f TFLOP/s (%4.0f W, %4.0f MHz, %.3f pJ/FLOP) | cuBLAS %6.3f ms %6.0f TFLOP/s (%4.0f W, %4.0f MHz, %.3f pJ/FLOP)"
            % (name, ms_o, fl / ms_o / 1e9, w_o, f_o, w_o * ms_o * 1e-3 / fl * 1e12, ms_c, fl / ms_c / 1e9, w_c, f_c, w_c * ms_c * 1e-3 / fl * 1e12))
    print(line, flush=True)
    lines.append(line)
    del x, W, out
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "gemm_vs_cublas.txt"), "w").write("\n".join(lines) + "\n")
