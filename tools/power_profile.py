"""Board power, SM clock and time per launch of each kernel of the reverse step when it runs ALONE in a loop for a few seconds
(nvidia-smi sampled every 100 ms), next to the whole step: which kernels pull the board to its power limit, and how
much energy (power x time) each contributes to one step.  Evidence for DESIGN.md section 5 "What bounds the step".
python tools/power_profile.py [--batch 256] [--seconds 3]  ->  gpurun_out/power_profile.txt"""
import argparse, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from musediffusion_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--seconds", type=float, default=3.0)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, L, H, F, NH = a.batch, 2096, 768, 3072, 12
M = B * L
bf = torch.bfloat16


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop = [], False

    def run(self):
        while not self.stop:
            out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=power.draw,clocks.sm,power.limit,clocks_event_reasons.sw_power_cap",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip().split(",")
            if len(out) >= 4:
                self.rows.append((float(out[0]), float(out[1]), float(out[2]), out[3].strip()))
            time.sleep(0.1)


def med(v):
    v = sorted(v)
    return v[len(v) // 2]


def measure(name, fn, launches_per_step):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = Sampler()
    s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, tic = 0, time.perf_counter()
    e0.record()
    while time.perf_counter() - tic < a.seconds:
        for _ in range(4):
            fn()
        n += 4
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    s.stop = True
    s.join()
    rows = s.rows[len(s.rows) // 3:]                      # drop the ramp
    ms = e0.elapsed_time(e1) / n
    w, mhz, lim = med([r[0] for r in rows]), med([r[1] for r in rows]), rows[0][2]
    capped = sum(r[3].lower().startswith("active") for r in rows) / len(rows)
    line = "%-34s %8.3f ms/launch  %6.0f W of %4.0f  %5.0f MHz  power-capped %3.0f%% of samples  -> %7.2f J per step (%d launches)" % (
        name, ms, w, lim, mhz, 100 * capped, w * ms * 1e-3 * launches_per_step, launches_per_step)
    print(line, flush=True)
    return line


x = (torch.randn(M, H, device=dev) * 0.5).to(bf)
xf = (torch.randn(M, F, device=dev) * 0.5).to(bf)
qkv = (torch.randn(M, 3 * H, device=dev) * 0.7).to(bf)
Wqkv = (torch.randn(3 * H, H, device=dev) * 0.03).to(bf)
Wo = (torch.randn(H, H, device=dev) * 0.03).to(bf)
W1 = (torch.randn(F, H, device=dev) * 0.03).to(bf)
W2 = (torch.randn(H, F, device=dev) * 0.03).to(bf)
b3, b1, bh = torch.zeros(3 * H, device=dev), torch.zeros(F, device=dev), torch.zeros(H, device=dev)
g, be = torch.ones(H, device=dev), torch.zeros(H, device=dev)
o_qkv, o_h, o_f = torch.empty(M, 3 * H, device=dev, dtype=bf), torch.empty(M, H, device=dev, dtype=bf), torch.empty(M, F, device=dev, dtype=bf)
lines = ["B = %d sequences x %d tokens; each kernel alone in a loop for %.1f s" % (B, L, a.seconds)]
lines.append(measure("attention", lambda: ops.attention(qkv, B, L, NH, out=o_h), 12))
lines.append(measure("QKV GEMM 768 -> 2304", lambda: ops.linear(x, Wqkv, b3, out=o_qkv), 12))
lines.append(measure("out-proj GEMM 768 -> 768", lambda: ops.linear(x, Wo, bh, out=o_h), 12))
lines.append(measure("FFN1 GEMM 768 -> 3072 + erf-GELU", lambda: ops.linear(x, W1, b1, epilogue=_lib.EPI_BIAS_GELU, out=o_f), 12))
lines.append(measure("FFN2 GEMM 3072 -> 768", lambda: ops.linear(xf, W2, bh, out=o_h), 12))
lines.append(measure("LayerNorm(x + residual)", lambda: ops.layernorm(x, g, be, 1e-12, resid=o_h, out=o_h), 25))
del x, xf, qkv, o_qkv, o_h, o_f
torch.cuda.empty_cache()

# the whole reverse step (captured graph, as the bench runs it)
from functools import partial
from types import SimpleNamespace
from musediffusion_b200.initialization import create_model_and_diffusion
from musediffusion_b200.rounding import denoised_fn_round
from musediffusion_b200.sample import build_model_emb
from musediffusion_b200.synthetic import make_synthetic_batch
T = 2000
targs = SimpleNamespace(hidden_dim=128, hidden_t_dim=128, vocab_size=729, seq_len=L, dropout=0.1, noise_schedule="sqrt",
                        diffusion_steps=T, timestep_respacing="", rescale_timesteps=True, predict_xstart=True)
torch.manual_seed(0)
model, diffusion = create_model_and_diffusion(targs)
model.eval().requires_grad_(False).to(dev)
emb = build_model_emb(model, dev)
c = make_synthetic_batch("modification", B, L, seed=105)
x_start = model.get_embeds(torch.from_numpy(c["input_ids"]).to(dev))
mask = torch.broadcast_to(torch.from_numpy(c["input_mask"]).to(dev).unsqueeze(-1), x_start.shape)
xT = diffusion.q_sample(x_start.unsqueeze(-1), torch.full((B, 1), T - 1, device=dev), mask=mask).squeeze(-1)
gen = diffusion._loop(_lib.STEP_DDPM, model, tuple(xT.shape), xT, True, partial(denoised_fn_round, emb, dist=None), None, dev, False,
                      1, 0, True, mask, x_start, 0.0, list(range(T))[::-1], want_aux=False)
for _ in range(4):
    next(gen)
lines.append(measure("whole reverse step (graph replay)", lambda: next(gen), 1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "power_profile.txt"), "w").write("\n".join(lines) + "\n")
