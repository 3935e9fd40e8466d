"""Runs a few reverse-diffusion steps of the bench workload (for ncu captures): python tools/profile_step.py --batch 64 --steps 2"""
import argparse
import os
import sys
from functools import partial
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from musediffusion_b200 import _lib  # noqa: E402
from musediffusion_b200.initialization import create_model_and_diffusion  # noqa: E402
from musediffusion_b200.rounding import denoised_fn_round  # noqa: E402
from musediffusion_b200.sample import build_model_emb  # noqa: E402
from musediffusion_b200.synthetic import make_synthetic_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--seq-len", type=int, default=2096)
a = ap.parse_args()
dev = torch.device("cuda:0")
T = 2000
targs = SimpleNamespace(hidden_dim=128, hidden_t_dim=128, vocab_size=729, seq_len=a.seq_len, dropout=0.1,
                        noise_schedule="sqrt", diffusion_steps=T, timestep_respacing="", rescale_timesteps=True,
                        predict_xstart=True)
torch.manual_seed(0)
model, diffusion = create_model_and_diffusion(targs)
diffusion.use_cuda_graph = False          # plain launches: ncu's -k / -s / -c then count kernels in program order
model.eval().requires_grad_(False).to(dev)
emb = build_model_emb(model, dev)
c = make_synthetic_batch("modification", a.batch, a.seq_len, seed=105)
ids = torch.from_numpy(c["input_ids"]).to(dev)
x_start = model.get_embeds(ids)
mask = torch.broadcast_to(torch.from_numpy(c["input_mask"]).to(dev).unsqueeze(-1), x_start.shape)
x = diffusion.q_sample(x_start.unsqueeze(-1), torch.full((a.batch, 1), T - 1, device=dev), mask=mask).squeeze(-1)
out = diffusion.p_sample_loop(model, tuple(x.shape), noise=x, denoised_fn=partial(denoised_fn_round, emb, dist=None),
                              model_kwargs={}, top_p=1, clamp_step=0, clamp_first=True, mask=mask, x_start=x_start,
                              t_enc=a.steps, only_last=True)
tok = model.decode_tokens(out[-1])
torch.cuda.synchronize()
print("ok", tuple(tok.shape))
