"""Upper bounds for fusion work: time of the reverse step with one kind of launch REMOVED (the results are wrong — this is a
timing experiment only).  Under the power cap the step time follows the step's energy, so the marginal cost of a kernel
inside the step is not its stand-alone duration.   python tools/skip_experiment.py [--batch 256] [--steps 10]"""
import argparse, os, sys
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
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=10)
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
ids = torch.from_numpy(c["input_ids"]).to(dev)
x_start = model.get_embeds(ids)
mask = torch.broadcast_to(torch.from_numpy(c["input_mask"]).to(dev).unsqueeze(-1), x_start.shape)
x = diffusion.q_sample(x_start.unsqueeze(-1), torch.full((B, 1), T - 1, device=dev), mask=mask).squeeze(-1)
fn = partial(denoised_fn_round, emb, dist=None)
real = {"layernorm": ops.layernorm, "attention": ops.attention, "linear": ops.linear}


def timed(tag):
    gen = diffusion._loop(_lib.STEP_DDPM, model, tuple(x.shape), x, True, fn, None, dev, False, 1, 0, True, mask, x_start, 0.0,
                          list(range(T))[::-1][:a.steps + 6], want_aux=False)
    for _ in range(5):
        next(gen)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        next(gen)
    e1.record()
    torch.cuda.synchronize()
    print("%-28s %.2f ms/step" % (tag, e0.elapsed_time(e1) / a.steps))


def skip_ln(xx, g, b_, eps, resid=None, out=None):
    return out if out is not None else xx


def skip_att(qkv, B_, L_, NH, out=None):
    return out


def skip_gelu_linear(A, W, bias, epilogue=_lib.EPI_BIAS, **kw):
    return real["linear"](A, W, bias, _lib.EPI_BIAS if epilogue == _lib.EPI_BIAS_GELU else epilogue, **kw)


for rep in range(2):
    timed("full step")
    ops.layernorm = skip_ln
    timed("without the 25 LayerNorms")
    ops.layernorm = real["layernorm"]
    ops.attention = skip_att
    timed("without the 12 attentions")
    ops.attention = real["attention"]
    ops.linear = skip_gelu_linear
    timed("FFN1 without GELU")
    ops.linear = real["linear"]
