"""BASELINE.json config 4: sampling-step sweep through the public API (sample_batch) to separate per-step kernel cost
from fixed overhead: time = a + b * steps.   python tools/step_sweep.py --batch 512 --steps 50 200 1000 2000   (2000 = the full DDPM chain)"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from types import SimpleNamespace
from musediffusion_b200.initialization import create_model_and_diffusion, seed_all
from musediffusion_b200.sample import build_model_emb, sample_batch
from musediffusion_b200.synthetic import make_synthetic_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--steps", type=int, nargs="+", default=[50, 200, 1000, 2000])
ap.add_argument("--mode", default="modification")
a = ap.parse_args()
dev = torch.device("cuda:0")
T, L = 2000, 2096
targs = SimpleNamespace(hidden_dim=128, hidden_t_dim=128, vocab_size=729, seq_len=L, dropout=0.1, noise_schedule="sqrt",
                        diffusion_steps=T, timestep_respacing="", rescale_timesteps=True, predict_xstart=True)
torch.manual_seed(0)
model, diffusion = create_model_and_diffusion(targs)
model.eval().requires_grad_(False).to(dev)
emb = build_model_emb(model, dev)
seed_all(105)
cond = {k: torch.from_numpy(v).pin_memory() for k, v in make_synthetic_batch(a.mode, a.batch, L, seed=105).items() if k != "length"}
sample_batch(model, diffusion, emb, cond, a.mode, 10, T, strength=1.0)      # warm-up (DDIM, 10 steps)
torch.cuda.synchronize()
rows = []
for st in a.steps:
    torch.cuda.synchronize()
    tic = time.perf_counter()
    tok = sample_batch(model, diffusion, emb, cond, a.mode, st, T, strength=1.0).cpu()
    torch.cuda.synchronize()
    rows.append((st, time.perf_counter() - tic))
    print("steps %5d (%s): %.3f s  -> %.4f sequences/s, %.2f ms/step" % (st, "DDPM" if st == T else "DDIM gap %d" % (T // st), rows[-1][1], a.batch / rows[-1][1], 1e3 * rows[-1][1] / st), flush=True)
x = np.array([r[0] for r in rows], dtype=np.float64); y = np.array([r[1] for r in rows])
b, c = np.polyfit(x, y, 1)
print(json.dumps({"config": "step sweep, batch %d, %s, DDIM" % (a.batch, a.mode), "fit": "time = a + b*steps",
                  "a_seconds": c, "b_seconds_per_step": b, "points": rows,
                  "extrapolated_full_chain_seconds": c + b * T, "extrapolated_sequences_per_s": a.batch / (c + b * T)}))
