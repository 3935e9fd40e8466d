"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.

A plain numpy (CPU, fp32 arithmetic / float64 tables) restatement of the MuseDiffusion
reverse-diffusion sampling path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the product
package (`musediffusion_b200`) never does and fails loudly when its CUDA library is missing.

Parity status: PINNED against the unmodified reference, executed in the build container
(`oracle/make_golden.py` imports `/root/reference` through `oracle/ref_shim.py` and writes
`tests/golden/*.npz`; `tests/test_oracle_golden.py` replays them).  The reference has no
tests / golden vectors of its own, so these generated fixtures plus the
known-answer values of SURVEY.md Appendix B are the pin.

Every function cites the reference file:line (relative to /root/reference/) it restates.
The third-party piece (HF `transformers==4.22.2` `BertEncoder`, not vendored in the reference;
call sites MuseDiffusion/models/network.py:9-10,44-46,74,151) is restated from its published
algorithm: 12 post-LN layers of {QKV Linear, softmax(QK^T/sqrt(64))V over 12 heads without mask,
Linear, LN(x+.), Linear 768->3072, erf-GELU, Linear 3072->768, LN(x+.)}, eps=1e-12.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

try:  # scipy is in the image; fall back to math.erf vectorised (slow) if it is not
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float64])

F32 = np.float32

# --------------------------------------------------------------------------------------
# 1. beta schedules                               MuseDiffusion/models/diffusion.py:22-118
# --------------------------------------------------------------------------------------


def betas_for_alpha_bar(num_diffusion_timesteps: int, alpha_bar: Callable[[float], float],
                        max_beta: float = 0.999) -> np.ndarray:
    """diffusion.py:101-118."""
    betas = []
    for i in range(num_diffusion_timesteps):
        t1 = i / num_diffusion_timesteps
        t2 = (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return np.array(betas)


def betas_for_alpha_bar_left(num_diffusion_timesteps: int, alpha_bar: Callable[[float], float],
                             max_beta: float = 0.999) -> np.ndarray:
    """diffusion.py:80-98."""
    betas = [min(1 - alpha_bar(0), max_beta)]
    for i in range(num_diffusion_timesteps - 1):
        t1 = i / num_diffusion_timesteps
        t2 = (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return np.array(betas)


def get_named_beta_schedule(schedule_name: str, T: int) -> np.ndarray:
    """diffusion.py:22-77 (all six named schedules)."""
    if schedule_name == "linear":
        scale = 1000 / T
        return np.linspace(scale * 0.0001, scale * 0.02, T, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(T, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    if schedule_name == "sqrt":
        return betas_for_alpha_bar(T, lambda t: 1 - np.sqrt(t + 0.0001))
    if schedule_name == "trunc_cos":
        return betas_for_alpha_bar_left(T, lambda t: np.cos((t + 0.1) / 1.1 * np.pi / 2) ** 2)
    if schedule_name == "trunc_lin":
        scale = 1000 / T
        return np.linspace(scale * 0.0001 + 0.01, scale * 0.02 + 0.01, T, dtype=np.float64)
    if schedule_name == "pw_lin":
        scale = 1000 / T
        beta_start = scale * 0.0001 + 0.01
        beta_mid = scale * 0.0001
        beta_end = scale * 0.02
        return np.concatenate([np.linspace(beta_start, beta_mid, 10, dtype=np.float64),
                               np.linspace(beta_mid, beta_end, T - 10, dtype=np.float64)])
    raise NotImplementedError("unknown beta schedule: {}".format(schedule_name))


def space_timesteps(num_timesteps: int, section_counts) -> set:
    """diffusion.py:920-969."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired_count = int(section_counts[len("ddim"):])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired_count:
                    return set(range(0, num_timesteps, i))
            raise ValueError("cannot create exactly {} steps with an integer stride".format(num_timesteps))
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start_idx = 0
    all_steps = []
    for i, section_count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < section_count:
            raise ValueError("cannot divide section of {0} steps into {1}".format(size, section_count))
        frac_stride = 1 if section_count <= 1 else (size - 1) / (section_count - 1)
        cur_idx = 0.0
        taken = []
        for _ in range(section_count):
            taken.append(start_idx + round(cur_idx))
            cur_idx += frac_stride
        all_steps += taken
        start_idx += size
    return set(all_steps)


# --------------------------------------------------------------------------------------
# 2. coefficient tables + respacing     diffusion.py:136-185 (tables), :981-996 (respacing)
# --------------------------------------------------------------------------------------


class Schedule:
    """float64 coefficient tables of GaussianDiffusion.__init__ after SpacedDiffusion respacing."""

    def __init__(self, betas: np.ndarray, use_timesteps=None, rescale_timesteps: bool = True,
                 predict_xstart: bool = True):
        base = np.array(betas, dtype=np.float64)
        self.original_num_steps = len(base)
        if use_timesteps is None:
            use_timesteps = set(range(len(base)))
        # SpacedDiffusion.__init__ (diffusion.py:981-996)
        base_ac = np.cumprod(1.0 - base, axis=0)
        last = 1.0
        new_betas = []
        self.timestep_map: List[int] = []
        for i, ac in enumerate(base_ac):
            if i in use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        betas = np.array(new_betas, dtype=np.float64)
        assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.rescale_timesteps = rescale_timesteps
        self.predict_xstart = predict_xstart
        self.num_timesteps = int(betas.shape[0])
        # GaussianDiffusion.__init__ (diffusion.py:154-183)
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(
            np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        # fixed-large model variance, p_mean_variance (diffusion.py:313-314)
        self.model_variance = np.append(self.posterior_variance[1], self.betas[1:])
        self.model_log_variance = np.log(self.model_variance)

    def model_timestep(self, t: np.ndarray) -> np.ndarray:
        """_WrappedModel.__call__ (diffusion.py:1027-1032): loop index -> value fed to the denoiser."""
        new_ts = np.asarray(self.timestep_map, dtype=np.int64)[np.asarray(t, dtype=np.int64)]
        if self.rescale_timesteps:
            return new_ts.astype(F32) * F32(1000.0 / self.original_num_steps)
        return new_ts


def make_schedule(noise_schedule="sqrt", diffusion_steps=2000, timestep_respacing="",
                  rescale_timesteps=True, predict_xstart=True) -> Schedule:
    """create_model_and_diffusion, diffusion half (MuseDiffusion/utils/initialization.py:123-134)."""
    betas = get_named_beta_schedule(noise_schedule, diffusion_steps)
    if not timestep_respacing:
        timestep_respacing = [diffusion_steps]
    return Schedule(betas, space_timesteps(diffusion_steps, timestep_respacing),
                    rescale_timesteps=rescale_timesteps, predict_xstart=predict_xstart)


def extract(arr: np.ndarray, t: np.ndarray, ndim: int) -> np.ndarray:
    """_extract_into_tensor (diffusion.py:904-917): float64 table -> fp32, gather by t, broadcast."""
    res = np.asarray(arr, dtype=np.float64).astype(F32)[np.asarray(t, dtype=np.int64)]
    while res.ndim < ndim:
        res = res[..., None]
    return res


# --------------------------------------------------------------------------------------
# 3. forward noising / posterior / rounding
# --------------------------------------------------------------------------------------


def q_sample(s: Schedule, x_start: np.ndarray, t: np.ndarray, noise: np.ndarray,
             mask: Optional[np.ndarray] = None) -> np.ndarray:
    """q_sample (diffusion.py:229-255).  `mask` already broadcast to x_start.shape (0 = keep x_start)."""
    x_start = x_start.astype(F32)
    x_t = (extract(s.sqrt_alphas_cumprod, t, x_start.ndim) * x_start
           + extract(s.sqrt_one_minus_alphas_cumprod, t, x_start.ndim) * noise.astype(F32))
    if mask is None:
        return x_t
    return np.where(mask == 0, x_start, x_t)


def q_posterior_mean(s: Schedule, x_start: np.ndarray, x_t: np.ndarray, t: np.ndarray) -> np.ndarray:
    """q_posterior_mean_variance (diffusion.py:257-278), mean only (variances unused when sampling)."""
    return (extract(s.posterior_mean_coef1, t, x_t.ndim) * x_start
            + extract(s.posterior_mean_coef2, t, x_t.ndim) * x_t)


def efficient_knn(E: np.ndarray, x: np.ndarray):
    """get_efficient_knn (MuseDiffusion/models/rounding.py:21-28).

    Returns (idx [M] int64, dist [V, M] fp32).  Ties resolve to the lowest vocabulary index
    (torch.max(dim=0) on CPU returns the first maximum)."""
    E = E.astype(F32)
    x = x.reshape(-1, x.shape[-1]).astype(F32)
    emb_norm = (E ** 2).sum(-1).reshape(-1, 1)
    arr_norm = (x ** 2).sum(-1).reshape(-1, 1)
    dist = emb_norm + arr_norm.T - F32(2.0) * (E @ x.T)
    dist = np.clip(dist, 0.0, np.inf).astype(F32)
    idx = np.argmax(-dist, axis=0)            # first maximum == lowest index on ties
    return idx.astype(np.int64), dist


def top2_margin(dist: np.ndarray) -> np.ndarray:
    """distance gap between the nearest and second-nearest embedding (parity-report helper)."""
    part = np.partition(dist, 1, axis=0)
    return (part[1] - part[0]).astype(F32)


def denoised_fn_round(E: np.ndarray, x: np.ndarray) -> np.ndarray:
    """denoised_fn_round (rounding.py:31-47) with dist=None: nearest embedding row, same shape as x."""
    idx, _ = efficient_knn(E, x)
    return E.astype(F32)[idx].reshape(x.shape)


def process_xstart(E: Optional[np.ndarray], x: np.ndarray, clip_denoised: bool) -> np.ndarray:
    """p_mean_variance.process_xstart (diffusion.py:319-325): round first, then clamp(-1, 1)."""
    if E is not None:
        x = denoised_fn_round(E, x)
    if clip_denoised:
        x = np.clip(x, -1.0, 1.0)
    return x.astype(F32)


def predict_xstart(s: Schedule, model_output: np.ndarray, x: np.ndarray, t: np.ndarray,
                   E: Optional[np.ndarray], clip_denoised: bool) -> np.ndarray:
    """p_mean_variance predict_xstart / eps branch (diffusion.py:327-333, 194-199)."""
    if s.predict_xstart:
        return process_xstart(E, model_output, clip_denoised)
    x0 = (extract(s.sqrt_recip_alphas_cumprod, t, x.ndim) * x
          - extract(s.sqrt_recipm1_alphas_cumprod, t, x.ndim) * model_output)
    return process_xstart(E, x0, clip_denoised)


def p_sample_step(s: Schedule, x: np.ndarray, t: np.ndarray, model_output: np.ndarray, noise: np.ndarray,
                  E: Optional[np.ndarray], clip_denoised: bool = True,
                  mask: Optional[np.ndarray] = None, x_start: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """p_sample after the model call (diffusion.py:349-404 + p_mean_variance :311-347)."""
    pred = predict_xstart(s, model_output, x, t, E, clip_denoised)
    mean = q_posterior_mean(s, pred, x, t)
    logvar = extract(s.model_log_variance, t, x.ndim)
    nonzero = (np.asarray(t) != 0).astype(F32).reshape(-1, *([1] * (x.ndim - 1)))
    sample = mean + nonzero * np.exp(F32(0.5) * logvar) * noise.astype(F32)
    if mask is not None:
        sample = np.where(mask == 0, x_start, sample)
    return {"sample": sample.astype(F32), "pred_xstart": pred, "greedy_mean": mean.astype(F32)}


def ddim_step(s: Schedule, x: np.ndarray, t: np.ndarray, model_output: np.ndarray, noise: np.ndarray,
              E: Optional[np.ndarray], clip_denoised: bool = True, eta: float = 0.0,
              mask: Optional[np.ndarray] = None, x_start: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """ddim_sample after the model call (diffusion.py:701-757, _predict_eps_from_xstart :201-205)."""
    pred = predict_xstart(s, model_output, x, t, E, clip_denoised)
    eps = ((extract(s.sqrt_recip_alphas_cumprod, t, x.ndim) * x - pred)
           / extract(s.sqrt_recipm1_alphas_cumprod, t, x.ndim))
    alpha_bar = extract(s.alphas_cumprod, t, x.ndim)
    alpha_bar_prev = extract(s.alphas_cumprod_prev, t, x.ndim)
    sigma = (F32(eta) * np.sqrt((1 - alpha_bar_prev) / (1 - alpha_bar))
             * np.sqrt(1 - alpha_bar / alpha_bar_prev)).astype(F32)
    mean_pred = pred * np.sqrt(alpha_bar_prev) + np.sqrt(1 - alpha_bar_prev - sigma ** 2) * eps
    nonzero = (np.asarray(t) != 0).astype(F32).reshape(-1, *([1] * (x.ndim - 1)))
    sample = mean_pred + nonzero * sigma * noise.astype(F32)
    if mask is not None:
        sample = np.where(mask == 0, x_start, sample)
    return {"sample": sample.astype(F32), "pred_xstart": pred}


# --------------------------------------------------------------------------------------
# 4. denoiser                                   MuseDiffusion/models/network.py:31-158
# --------------------------------------------------------------------------------------


def timestep_embedding(timesteps: np.ndarray, dim: int, max_period: int = 10000) -> np.ndarray:
    """network.py:108-129."""
    half = dim // 2
    freqs = np.exp(F32(-math.log(max_period)) * np.arange(half, dtype=F32) / F32(half)).astype(F32)
    args = np.asarray(timesteps, dtype=F32)[:, None] * freqs[None]
    emb = np.concatenate([np.cos(args), np.sin(args)], axis=-1).astype(F32)
    if dim % 2:
        emb = np.concatenate([emb, np.zeros_like(emb[:, :1])], axis=-1)
    return emb


def _linear(x, w, b):
    return x @ w.T + b


def _layer_norm(x, g, b, eps=1e-12):
    mu = x.mean(-1, keepdims=True, dtype=F32)
    var = ((x - mu) ** 2).mean(-1, keepdims=True, dtype=F32)
    return ((x - mu) / np.sqrt(var + F32(eps)) * g + b).astype(F32)


def _gelu_erf(x):
    return (x * 0.5 * (1.0 + _erf(x / math.sqrt(2.0)))).astype(F32)


def _silu(x):
    return (x / (1.0 + np.exp(-x))).astype(F32)


def bert_layer(p: Dict[str, np.ndarray], pre: str, h: np.ndarray, num_heads: int) -> np.ndarray:
    """One HF BertLayer (transformers 4.22.2 modeling_bert.py BertSelfAttention/BertSelfOutput/
    BertIntermediate/BertOutput), hidden states only — no attention mask, no head mask, eval mode."""
    B, L, H = h.shape
    dh = H // num_heads
    a = pre + "attention."
    q = _linear(h, p[a + "self.query.weight"], p[a + "self.query.bias"])
    k = _linear(h, p[a + "self.key.weight"], p[a + "self.key.bias"])
    v = _linear(h, p[a + "self.value.weight"], p[a + "self.value.bias"])
    ctx = np.empty_like(q)
    scale = F32(1.0 / math.sqrt(dh))
    for b in range(B):                       # per (b, head): keeps the L x L scores small
        for hd in range(num_heads):
            sl = slice(hd * dh, (hd + 1) * dh)
            sc = (q[b, :, sl] @ k[b, :, sl].T) * scale
            sc = sc - sc.max(-1, keepdims=True)
            pr = np.exp(sc)
            pr /= pr.sum(-1, keepdims=True)
            ctx[b, :, sl] = pr @ v[b, :, sl]
    so = _linear(ctx, p[a + "output.dense.weight"], p[a + "output.dense.bias"])
    h1 = _layer_norm(so + h, p[a + "output.LayerNorm.weight"], p[a + "output.LayerNorm.bias"])
    inter = _gelu_erf(_linear(h1, p[pre + "intermediate.dense.weight"], p[pre + "intermediate.dense.bias"]))
    out = _linear(inter, p[pre + "output.dense.weight"], p[pre + "output.dense.bias"])
    return _layer_norm(out + h1, p[pre + "output.LayerNorm.weight"], p[pre + "output.LayerNorm.bias"])


def denoiser_forward(p: Dict[str, np.ndarray], x: np.ndarray, timesteps: np.ndarray,
                     num_heads: int = 12, hidden_t_dim: int = None) -> np.ndarray:
    """TransformerNetModel.forward (network.py:131-158); `p` uses the reference state-dict keys.  hidden_t_dim defaults to
    the input width of time_embed.0 (network.py:57-61: the sinusoid is hidden_t_dim wide)."""
    if hidden_t_dim is None:
        hidden_t_dim = p["time_embed.0.weight"].shape[1]
    x = x.astype(F32)
    B, L, _ = x.shape
    emb_t = _linear(_silu(_linear(timestep_embedding(timesteps, hidden_t_dim),
                                  p["time_embed.0.weight"], p["time_embed.0.bias"])),
                    p["time_embed.2.weight"], p["time_embed.2.bias"])
    if "input_up_proj.0.weight" in p:
        emb_x = _linear(np.tanh(_linear(x, p["input_up_proj.0.weight"], p["input_up_proj.0.bias"])),
                        p["input_up_proj.2.weight"], p["input_up_proj.2.bias"])
    else:
        emb_x = x
    h = p["position_embeddings.weight"][:L][None] + emb_x + emb_t[:, None, :]
    h = _layer_norm(h.astype(F32), p["LayerNorm.weight"], p["LayerNorm.bias"])
    n_layers = 1 + max(int(k.split(".")[2]) for k in p if k.startswith("input_transformers.layer."))
    for i in range(n_layers):
        h = bert_layer(p, "input_transformers.layer.%d." % i, h, num_heads)
    if "output_down_proj.0.weight" in p:
        h = _linear(np.tanh(_linear(h, p["output_down_proj.0.weight"], p["output_down_proj.0.bias"])),
                    p["output_down_proj.2.weight"], p["output_down_proj.2.bias"])
    return h.astype(F32)


def get_embeds(p: Dict[str, np.ndarray], input_ids: np.ndarray) -> np.ndarray:
    """network.py:88-89."""
    return p["word_embedding.weight"].astype(F32)[np.asarray(input_ids, dtype=np.int64)]


def get_logits(p: Dict[str, np.ndarray], hidden: np.ndarray) -> np.ndarray:
    """network.py:91-93 (logits_mode=1): lm_head(x) = x E^T + b, weight tied to word_embedding (:55-58)."""
    return (hidden.astype(F32) @ p["word_embedding.weight"].astype(F32).T + p["lm_head.bias"]).astype(F32)


def logits_argmax(p: Dict[str, np.ndarray], hidden: np.ndarray) -> np.ndarray:
    """MuseDiffusion/run/sample.py:219-220."""
    return np.argmax(get_logits(p, hidden), axis=-1).astype(np.int64)


# --------------------------------------------------------------------------------------
# 5. parameter construction (deterministic random init shared by oracle / reference / CUDA)
# --------------------------------------------------------------------------------------


def make_random_params(seed: int = 0, seq_len: int = 2096, vocab_size: int = 729, hidden_dim: int = 128,
                       hidden_t_dim: int = 128, hidden: int = 768, ffn: int = 3072, layers: int = 12,
                       emb_scale: float = 1.0) -> Dict[str, np.ndarray]:
    """State dict with the reference's 211 keys/shapes, filled from numpy PCG64.

    Scales follow torch defaults (nn.Linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in)); nn.Embedding N(0,1);
    LayerNorm weight 1 / bias 0 perturbed slightly so the affine part is exercised)."""
    rng = np.random.default_rng(seed)
    p: Dict[str, np.ndarray] = {}

    def lin(name, out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        p[name + ".weight"] = rng.uniform(-bound, bound, size=(out_f, in_f)).astype(F32)
        p[name + ".bias"] = rng.uniform(-bound, bound, size=(out_f,)).astype(F32)

    def ln(name, n):
        p[name + ".weight"] = (1.0 + 0.05 * rng.standard_normal(n)).astype(F32)
        p[name + ".bias"] = (0.05 * rng.standard_normal(n)).astype(F32)

    p["word_embedding.weight"] = (emb_scale * rng.standard_normal((vocab_size, hidden_dim))).astype(F32)
    p["lm_head.weight"] = p["word_embedding.weight"]
    p["lm_head.bias"] = rng.uniform(-1 / math.sqrt(hidden_dim), 1 / math.sqrt(hidden_dim), vocab_size).astype(F32)
    lin("time_embed.0", hidden_t_dim * 4, hidden_t_dim)
    lin("time_embed.2", hidden, hidden_t_dim * 4)
    lin("input_up_proj.0", hidden, hidden_dim)
    lin("input_up_proj.2", hidden, hidden)
    p["position_ids"] = np.arange(seq_len, dtype=np.int64)[None]
    p["position_embeddings.weight"] = rng.standard_normal((seq_len, hidden)).astype(F32)
    ln("LayerNorm", hidden)
    for i in range(layers):
        pre = "input_transformers.layer.%d." % i
        lin(pre + "attention.self.query", hidden, hidden)
        lin(pre + "attention.self.key", hidden, hidden)
        lin(pre + "attention.self.value", hidden, hidden)
        lin(pre + "attention.output.dense", hidden, hidden)
        ln(pre + "attention.output.LayerNorm", hidden)
        lin(pre + "intermediate.dense", ffn, hidden)
        lin(pre + "output.dense", hidden, ffn)
        ln(pre + "output.LayerNorm", hidden)
    lin("output_down_proj.0", hidden, hidden)
    lin("output_down_proj.2", hidden_dim, hidden)
    return p


# --------------------------------------------------------------------------------------
# 6. synthetic ComMU-shaped inputs                      SURVEY.md §8d; formats §8a-19
# --------------------------------------------------------------------------------------

_META_RANGES = [(560, 600), (601, 625), (626, 629), (630, 637), (638, 640), (641, 649),
                (650, 652), (653, 718), (653, 718), (719, 725), (726, 728)]
# commu/preprocessor/encoder/event_tokens.py:308-329 (TOKEN_OFFSET), one token per meta field


def make_prefix(rng: np.random.Generator) -> List[int]:
    """11 meta tokens + chord tokens as MetaToSequence builds them (utils/decode_util.py:25-46)."""
    toks = [int(rng.integers(lo, hi + 1)) for lo, hi in _META_RANGES]
    n_bars = int(rng.choice([4, 8, 16]))
    for _ in range(n_bars):
        toks += [432, int(rng.integers(195, 304))]
        if rng.random() < 0.25:
            toks += [432 + 16 * int(rng.integers(1, 8)), int(rng.integers(195, 304))]
    return toks


def make_synthetic_batch(mode: str, B: int, L: int, seed: int = 105, per_row_prefix: bool = False):
    """generation: meta_to_batch format (utils/decode_util.py:221-230), int32.
    modification: collate_batches format (data/wrapper.py:90-126, data/preprocess.py:50-56), int64."""
    rng = np.random.default_rng(seed)
    if mode == "generation":
        ids = np.zeros((B, L), dtype=np.int32)
        msk = np.ones((B, L), dtype=np.int32)
        prefix = make_prefix(rng)
        for b in range(B):
            if per_row_prefix and b:
                prefix = make_prefix(rng)
            n = min(len(prefix), L - 1)
            ids[b, :n] = prefix[:n]
            msk[b, :n + 1] = 0
        return {"input_ids": ids, "input_mask": msk}
    if mode == "modification":
        ids = np.zeros((B, L), dtype=np.int64)
        msk = np.ones((B, L), dtype=np.int64)
        length = np.zeros((B,), dtype=np.int64)
        for b in range(B):
            prefix = make_prefix(rng)
            n = min(len(prefix), max(L // 2 - 1, 1))
            total = int(rng.integers(min(64, L), L + 1))
            row = list(prefix[:n]) + [1]
            k = 0
            while len(row) < total - 1:
                if k % 6 == 0:
                    row.append(2)
                row += [int(rng.integers(432, 560)), int(rng.integers(131, 195)),
                        int(rng.integers(3, 131)), int(rng.integers(304, 432))]
                k += 1
            row = row[:total - 1] + [1]
            ids[b, :len(row)] = row
            msk[b, :n + 1] = 0
            length[b] = len(row)
        return {"input_ids": ids, "input_mask": msk, "length": length}
    raise ValueError(mode)


# --------------------------------------------------------------------------------------
# 7. the loops                  diffusion.py:406-540 (DDPM), :797-901 (DDIM); run/sample.py:177-220
# --------------------------------------------------------------------------------------


class NoiseStream:
    """Mirrors the reference's calls to torch.randn_like with a numpy PCG64 stream so oracle, patched
    reference and CUDA path can all consume identical noise.  `randn(shape)` is one randn_like call."""

    def __init__(self, seed: int):
        self.rng = np.random.default_rng(seed)

    def randn(self, shape) -> np.ndarray:
        return self.rng.standard_normal(size=shape, dtype=np.float32)

    def truncated(self, shape, top_p) -> np.ndarray:
        """p_sample noise (diffusion.py:376-388): rejection-resample entries with |n| > top_p."""
        noise = self.randn(shape)
        if top_p is not None and top_p > 0:
            replace = np.abs(noise) > top_p
            while replace.any():
                noise[replace] = self.randn((int(replace.sum()),))
                replace = np.abs(noise) > top_p
        return noise


def ddpm_indices(s: Schedule, t_enc=None) -> List[int]:
    """diffusion.py:508."""
    return list(range(s.num_timesteps))[::-1][slice(t_enc)]


def ddim_indices(s: Schedule, gap: int = 1, t_enc=None) -> List[int]:
    """diffusion.py:878."""
    return list(range(s.num_timesteps))[::-1][::gap][slice(t_enc)]


def p_sample_loop(s: Schedule, p: Dict[str, np.ndarray], x_noised: np.ndarray, noise: NoiseStream,
                  clip_denoised=True, round_to_emb=True, top_p=1, clamp_step=0, clamp_first=True,
                  mask=None, x_start=None, t_enc=None, record: Optional[list] = None,
                  model_fn=None) -> np.ndarray:
    """p_sample_loop / p_sample_loop_progressive (diffusion.py:406-540), only_last semantics."""
    x = x_noised.astype(F32)
    E = p["word_embedding.weight"]
    B = x.shape[0]
    for i in ddpm_indices(s, t_enc):
        t = np.full((B,), i, dtype=np.int64)
        if not clamp_first:
            use_round = not (i > clamp_step)
        else:
            use_round = i >= clamp_step
        mo = (model_fn or denoiser_forward)(p, x, s.model_timestep(t))
        n = noise.truncated(x.shape, top_p)
        out = p_sample_step(s, x, t, mo, n, E if (round_to_emb and use_round) else None,
                            clip_denoised, mask, x_start)
        if record is not None:
            record.append({"t": i, "x_t": x, "model_output": mo, "noise": n, "sample": out["sample"]})
        x = out["sample"]
    return x


def ddim_sample_loop(s: Schedule, p: Dict[str, np.ndarray], x_noised: np.ndarray, noise: NoiseStream,
                     clip_denoised=True, round_to_emb=True, mask=None, x_start=None, gap=1, eta=0.0,
                     t_enc=None, record: Optional[list] = None, model_fn=None) -> np.ndarray:
    """ddim_sample_loop / _progressive (diffusion.py:797-901): always rounds, ignores top_p/clamp_*."""
    x = x_noised.astype(F32)
    E = p["word_embedding.weight"]
    B = x.shape[0]
    for i in ddim_indices(s, gap, t_enc):
        t = np.full((B,), i, dtype=np.int64)
        mo = (model_fn or denoiser_forward)(p, x, s.model_timestep(t))
        n = noise.randn(x.shape)                      # drawn even though eta = 0 (diffusion.py:738)
        out = ddim_step(s, x, t, mo, n, E if round_to_emb else None, clip_denoised, eta, mask, x_start)
        if record is not None:
            record.append({"t": i, "x_t": x, "model_output": mo, "noise": n, "sample": out["sample"]})
        x = out["sample"]
    return x


def sample_batch(s: Schedule, p: Dict[str, np.ndarray], cond: Dict[str, np.ndarray], mode: str,
                 step: int, noise: NoiseStream, strength: float = 0.75, top_p=1, clamp_step=0,
                 clip_denoised=True, record: Optional[list] = None, model_fn=None) -> np.ndarray:
    """The hot slice of run/sample.py:177-220 for one batch -> int64 tokens [B, L]."""
    ids = np.asarray(cond["input_ids"])
    x_start = get_embeds(p, ids)
    mask = np.broadcast_to(np.asarray(cond["input_mask"])[..., None], x_start.shape)
    diffusion_steps = s.original_num_steps
    if mode == "generation":
        noising_t = None
        x_noised = np.where(mask == 0, x_start, noise.randn(x_start.shape))        # sample.py:190-193
    else:
        noising_t = int(step * strength)                                            # sample.py:195
        t = np.full((ids.shape[0], 1), noising_t - 1, dtype=np.int64)
        x_noised = q_sample(s, x_start[..., None], t, noise.randn(x_start.shape + (1,)),
                            mask=mask[..., None])[..., 0]                           # sample.py:196-197
    if step == diffusion_steps:                                                     # sample.py:109-114
        out = p_sample_loop(s, p, x_noised, noise, clip_denoised, True, top_p, clamp_step, True,
                            mask, x_start, noising_t, record, model_fn)
    else:
        out = ddim_sample_loop(s, p, x_noised, noise, clip_denoised, True, mask, x_start,
                               diffusion_steps // step, 0.0, noising_t, record, model_fn)
    return logits_argmax(p, out)
