"""SM clock during the bench step, resolved in time: a one-warp probe kernel samples (globaltimer, clock64) every 100 us on
a side stream while K reverse steps run; kernel boundaries come from CUDA events converted to the same time base.
python tools/clock_timeline.py [--batch 256] [--steps 3]  ->  gpurun_out/clock_timeline.txt"""
import argparse, ctypes, os, sys
from functools import partial
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from musediffusion_b200 import _lib, ops
from musediffusion_b200.initialization import create_model_and_diffusion
from musediffusion_b200.rounding import denoised_fn_round
from musediffusion_b200.sample import build_model_emb
from musediffusion_b200.synthetic import make_synthetic_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--interval-us", type=int, default=100)
a = ap.parse_args()
probe = ctypes.CDLL(os.path.join(ROOT, "tools", "microbench", "libclockprobe.so"))
probe.clock_probe_launch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p]
dev = torch.device("cuda:0")
T, L = 2000, 2096
targs = SimpleNamespace(hidden_dim=128, hidden_t_dim=128, vocab_size=729, seq_len=L, dropout=0.1, noise_schedule="sqrt",
                        diffusion_steps=T, timestep_respacing="", rescale_timesteps=True, predict_xstart=True)
torch.manual_seed(0)
model, diffusion = create_model_and_diffusion(targs)
model.eval().requires_grad_(False).to(dev)
emb = build_model_emb(model, dev)
c = make_synthetic_batch("modification", a.batch, L, seed=105)
ids = torch.from_numpy(c["input_ids"]).to(dev)
x_start = model.get_embeds(ids)
mask = torch.broadcast_to(torch.from_numpy(c["input_mask"]).to(dev).unsqueeze(-1), x_start.shape)
x = diffusion.q_sample(x_start.unsqueeze(-1), torch.full((a.batch, 1), T - 1, device=dev), mask=mask).squeeze(-1)
fn = partial(denoised_fn_round, emb, dist=None)
warm = 4
gen = diffusion._loop(_lib.STEP_DDPM, model, tuple(x.shape), x, True, fn, None, dev, False, 1, 0, True, mask, x_start, 0.0,
                      list(range(T))[::-1][:warm + a.steps + 1], want_aux=False)
for _ in range(warm):
    next(gen)
torch.cuda.synchronize()
n = int((a.steps * 0.2 + 0.1) * 1e6 / a.interval_us)
buf = torch.zeros(2 * n, dtype=torch.int64, device=dev)
side = torch.cuda.Stream()
probe.clock_probe_launch(buf.data_ptr(), n, a.interval_us * 1000, side.cuda_stream)
import time
time.sleep(0.02)
prof = ops.profile_step(lambda: [next(gen) for _ in range(a.steps)])
torch.cuda.synchronize()
s = buf.cpu().numpy().reshape(n, 2)
s = s[s[:, 0] > 0]
t = (s[:, 0] - s[0, 0]) * 1e-6            # ms
f = np.diff(s[:, 1]) / np.diff(s[:, 0]) * 1e3      # MHz
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "clock_timeline.txt"), "w") as fo:
    fo.write("# kernel durations of the profiled steps (ms), in launch order\n")
    acc = 0.0
    for name, detail, ms in prof:
        fo.write("%8.3f %8.3f %s %s\n" % (acc, ms, name, detail))
        acc += ms
    fo.write("# t_ms  sm_mhz (probe, %d us windows)\n" % a.interval_us)
    for ti, fi in zip(t[1:], f):
        fo.write("%9.3f %7.1f\n" % (ti, fi))
print("samples", len(f), "clock MHz: min %.0f median %.0f max %.0f" % (f.min(), np.median(f), f.max()))
# coarse histogram of the clock over the busy part
busy = f[(t[1:] > 25)]
print("percentiles 5/25/50/75/95:", np.percentile(busy, [5, 25, 50, 75, 95]).round(0))
