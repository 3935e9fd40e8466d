"""Per-call (not per-step) time of sample_batch() at the bench size: calls of 6 / 12 / 24 chain steps give the slope and the
intercept of the call-time line.  With the `torch.cuda.graph` context the intercept jumped between 6 and 230 ms (the
context empties the caching allocator on entry: ~10 GB of step activations returned and re-acquired per call); with the
plain capture into a pool that outlives the graphs (diffusion.py::_capture) it is 8 ms on every call.
python tools/e2e_overhead.py [--batch 256]"""
import argparse, os, sys, time
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from musediffusion_b200.initialization import create_model_and_diffusion
from musediffusion_b200.sample import build_model_emb, sample_batch
from musediffusion_b200.synthetic import make_synthetic_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
a = ap.parse_args()
dev = torch.device("cuda:0")
T, L, B = 2000, 2096, a.batch
targs = SimpleNamespace(hidden_dim=128, hidden_t_dim=128, vocab_size=729, seq_len=L, dropout=0.1, noise_schedule="sqrt",
                        diffusion_steps=T, timestep_respacing="", rescale_timesteps=True, predict_xstart=True)
torch.manual_seed(0)
model, diffusion = create_model_and_diffusion(targs)
model.eval().requires_grad_(False).to(dev)
emb = build_model_emb(model, dev)
c = make_synthetic_batch("modification", B, L, seed=105)
cond = {k: torch.from_numpy(v).pin_memory() for k, v in c.items()}


def call(k):
    torch.cuda.synchronize()
    tic = time.perf_counter()
    tok = sample_batch(model, diffusion, emb, cond, "modification", T, T, strength=k / T, top_p=1, clamp_step=0, device=dev)
    host = tok.cpu()
    torch.cuda.synchronize()
    return time.perf_counter() - tic


call(6)
res = {}
for k in (6, 12, 24, 6, 12, 24):
    res.setdefault(k, []).append(call(k))
    print("k=%2d  %.4f s   reserved %.1f GB" % (k, res[k][-1], torch.cuda.memory_reserved() / 1e9), flush=True)
t6, t12, t24 = (min(res[k]) for k in (6, 12, 24))
slope = (t24 - t6) / 18
print("per step %.2f ms, per-call intercept %.1f ms (from k=6,24); check with k=12: predicted %.4f measured %.4f"
      % (slope * 1e3, (t6 - 6 * slope) * 1e3, t6 + 6 * slope, t12))
