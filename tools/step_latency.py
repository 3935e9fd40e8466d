"""Launch overhead at the reference's own operating points (BASELINE.json configs[0]: batch 4; CLI default batch 50):
wall time per reverse step of the eager loop (ops dispatched one by one from Python) and of the captured-graph loop,
next to the sum of the kernel times of one step (CUDA events around every launch).
python tools/step_latency.py [--batch 4] [--steps 60]"""
import argparse, os, sys, time
from functools import partial
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from musediffusion_b200 import _lib, ops
from musediffusion_b200.initialization import create_model_and_diffusion
from musediffusion_b200.rounding import denoised_fn_round
from musediffusion_b200.sample import build_model_emb
from musediffusion_b200.synthetic import make_synthetic_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--seq-len", type=int, default=2096)
a = ap.parse_args()
dev = torch.device("cuda:0")
T, L, B = 2000, a.seq_len, a.batch
targs = SimpleNamespace(hidden_dim=128, hidden_t_dim=128, vocab_size=729, seq_len=L, dropout=0.1, noise_schedule="sqrt",
                        diffusion_steps=T, timestep_respacing="", rescale_timesteps=True, predict_xstart=True)
torch.manual_seed(0)
model, diffusion = create_model_and_diffusion(targs)
model.eval().requires_grad_(False).to(dev)
emb = build_model_emb(model, dev)
c = make_synthetic_batch("modification", B, L, seed=105)
ids = torch.from_numpy(c["input_ids"]).to(dev)
x_start = model.get_embeds(ids)
mask = torch.broadcast_to(torch.from_numpy(c["input_mask"]).to(dev).unsqueeze(-1), x_start.shape)
x = diffusion.q_sample(x_start.unsqueeze(-1), torch.full((B, 1), T - 1, device=dev), mask=mask).squeeze(-1)
fn = partial(denoised_fn_round, emb, dist=None)


def run(use_graph, n):
    diffusion.use_cuda_graph = use_graph
    gen = diffusion._loop(_lib.STEP_DDPM, model, tuple(x.shape), x, True, fn, None, dev, False, 1, 0, True, mask, x_start, 0.0,
                          list(range(T))[::-1][:n + 8], want_aux=False)
    for _ in range(6):
        next(gen)
    torch.cuda.synchronize()
    tic = time.perf_counter()
    for _ in range(n):
        next(gen)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - tic) / n * 1e3
    return wall, gen


wall_eager, gen = run(False, a.steps)
prof = ops.profile_step(lambda: next(gen))
kernel_sum = sum(ms for _, _, ms in prof)
wall_graph, _ = run(True, a.steps)
print("B=%d L=%d: %d launches/step; sum of kernel times %.3f ms; wall per step eager %.3f ms (%.2fx), graph %.3f ms (%.2fx)"
      % (B, L, len(prof), kernel_sum, wall_eager, wall_eager / kernel_sum, wall_graph, wall_graph / kernel_sum))
