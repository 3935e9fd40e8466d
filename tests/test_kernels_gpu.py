"""GPU parity tests of the individual CUDA kernels, called through the C-ABI (ctypes) wrappers in
musediffusion_b200.ops.  Dense bf16 kernels are checked against a torch fp32 evaluation of the same op on the same
bf16-rounded operands (tolerance = bf16 output rounding); fp32 / integer kernels against the numpy oracle."""
import math

import numpy as np
import pytest
import torch

import musediff_oracle as O

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from musediffusion_b200 import _lib, ops  # noqa: E402

DEV = torch.device("cuda:0")


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


# ------------------------------------------------------------------------------------------------ linear
def _gelu_erf(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 768, 128), (300, 768, 768), (2096, 2304, 768),
                                   (2096 * 2, 3072, 768), (1000, 768, 3072), (2096, 128, 768), (520, 1024, 1024), (77, 200, 72), (1500, 512, 256), (4000, 768, 768)])
@pytest.mark.parametrize("epi", [_lib.EPI_BIAS, _lib.EPI_BIAS_GELU, _lib.EPI_BIAS_TANH])
def test_linear(M, N, K, epi):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + epi)
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    ref = A.float() @ W.float().T + bias
    if epi == _lib.EPI_BIAS_GELU:
        ref = _gelu_erf(ref)
    elif epi == _lib.EPI_BIAS_TANH:
        ref = torch.tanh(ref)
    out = ops.linear(A, W, bias, epi)
    torch.cuda.synchronize()
    assert out.dtype == torch.bfloat16 and out.shape == (M, N)
    assert rel_err(out, ref) < 1.2e-2, rel_err(out, ref)
    out32 = ops.linear(A, W, bias, epi, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_err(out32, ref) < 2e-3, rel_err(out32, ref)


def test_linear_pos_time():
    B, L, N, K = 3, 200, 768, 768
    g = torch.Generator(device="cpu").manual_seed(5)
    A = (torch.randn(B * L, K, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    pos = torch.randn(L, N, generator=g).to(DEV)
    temb = torch.randn(B, N, generator=g).to(DEV)
    ref = (A.float() @ W.float().T + bias).view(B, L, N) + pos[None] + temb[:, None]
    out = ops.linear(A, W, bias, _lib.EPI_BIAS_POS_TIME, out_dtype=torch.float32, pos=pos, temb=temb, temb_stride=N, L=L)
    assert rel_err(out.view(B, L, N), ref) < 2e-3
    ref1 = (A.float() @ W.float().T + bias).view(B, L, N) + pos[None] + temb[:1, None]
    out1 = ops.linear(A, W, bias, _lib.EPI_BIAS_POS_TIME, out_dtype=torch.float32, pos=pos, temb=temb, temb_stride=0, L=L)
    assert rel_err(out1.view(B, L, N), ref1) < 2e-3


def test_linear_argument_errors():
    A = torch.zeros(8, 12, dtype=torch.bfloat16, device=DEV)
    W = torch.zeros(16, 12, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(_lib.MuseDiffLibraryError):
        ops.linear(A, W, None)          # K not a multiple of 8
    with pytest.raises(_lib.MuseDiffLibraryError):
        ops.linear(torch.zeros(8, 16, dtype=torch.bfloat16), torch.zeros(16, 16, dtype=torch.bfloat16), None)  # CPU


# ------------------------------------------------------------------------------------------------ attention
# L picks the work-item kinds and tail paths of the kernel: 64 / 128 SINGLE (one Q tile, one key block; half-width /
# full last block), 130 / 200 / 256 PAIR only (last key block with 2 / 72 / 128 keys), 300 / 320 / 385 PAIR + SPLIT with a
# half-width tail (44 / 64 / 1 keys), 2048 all PAIR no tail, 2096 the base shape (8 PAIR + SPLIT, 48-key tail, 48-row last
# Q tile), 2176 = 17 x 128 SPLIT with a full tail, 4192 the scaled config (SPLIT, 96-key masked full-width tail)
@pytest.mark.parametrize("B,L,NH", [(1, 64, 2), (2, 128, 12), (2, 130, 3), (2, 200, 12), (1, 256, 3), (3, 300, 4), (2, 320, 2),
                                    (2, 385, 5), (2, 2048, 3), (2, 2096, 12), (1, 2176, 4), (1, 4192, 2)])
def test_attention(B, L, NH):
    H = NH * 64
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + L + NH)
    qkv = torch.randn(B * L, 3 * H, generator=g)
    qkv[:, :H] *= 0.5           # q (already carries the 1/sqrt(64) scale in the real pipeline)
    qkv[:, H:2 * H] *= 1.5
    qkv = qkv.to(torch.bfloat16).to(DEV)
    out = ops.attention(qkv, B, L, NH)
    torch.cuda.synchronize()
    q, k, v = [t.float().view(B, L, NH, 64).transpose(1, 2) for t in qkv.split(H, dim=1)]
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    ref = (p @ v).transpose(1, 2).reshape(B * L, H)
    assert rel_err(out, ref) < 1.5e-2, rel_err(out, ref)


def test_attention_peaky_rows_trigger_rescale():
    """keys whose score grows along the sequence: the running max rises by far more than the lazy-rescale threshold
    (2^32) from block to block, so the O / l rescale in TMEM runs repeatedly."""
    B, L, NH = 1, 640, 1
    g = torch.Generator(device="cpu").manual_seed(9)
    q = torch.randn(L, 64, generator=g)
    k = torch.randn(L, 64, generator=g) * 0.1
    k += (torch.arange(L)[:, None] / L) * q.mean(0, keepdim=True).sign() * 0.0
    k[:, 0] += torch.linspace(0, 40, L)         # later keys score much higher for rows with q[:,0] > 0
    q[:, 0] = q[:, 0].abs() * 3
    v = torch.randn(L, 64, generator=g)
    qkv = torch.cat([q, k, v], dim=1).to(torch.bfloat16).to(DEV)
    out = ops.attention(qkv, B, L, NH)
    qf, kf, vf = [t.float() for t in qkv.split(64, dim=1)]
    ref = torch.softmax(qf @ kf.T, dim=-1) @ vf
    assert rel_err(out, ref) < 1.5e-2, rel_err(out, ref)


def test_attention_many_work_items_deterministic():
    """Every CTA walks ~18 work items (Q double buffering, one-wait-per-block hand-offs, TMA-store epilogue reuse):
    two runs must agree bit for bit (a hand-off race shows up as run-to-run noise) and match fp32 softmax."""
    B, L, NH = 24, 2096, 12
    H = NH * 64
    g = torch.Generator(device="cpu").manual_seed(77)
    qkv = torch.randn(B * L, 3 * H, generator=g)
    qkv[:, :H] *= 0.5
    qkv = qkv.to(torch.bfloat16).to(DEV)
    out1 = ops.attention(qkv, B, L, NH).clone()
    out2 = ops.attention(qkv, B, L, NH).clone()
    torch.cuda.synchronize()
    assert torch.equal(out1, out2)
    worst = 0.0
    for b in (0, 11, 23):                      # fp32 reference for three of the sequences
        rows = slice(b * L, (b + 1) * L)
        q, k, v = [t.float().view(L, NH, 64).transpose(0, 1) for t in qkv[rows].split(H, dim=1)]
        ref = (torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v).transpose(0, 1).reshape(L, H)
        worst = max(worst, rel_err(out1[rows], ref))
    assert worst < 1.5e-2, worst


def test_split_embedding_cache_survives_address_reuse():
    """A freed embedding table's address is handed to the next table of the same shape by the caching allocator: the
    cached split must never be served for different contents (regression: stale split -> every free token rounded
    against another model's embeddings)."""
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(500, 128, generator=g).to(DEV)
    for k in range(6):
        E = (torch.randn(729, 128, generator=g) * (1 + k)).to(DEV)
        want = ops.round_argmin(x, E)
        got = ops.round_argmin_tc(x, ops.split_embedding(E))
        assert torch.equal(got, want), k
        del E
        torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------ layernorm etc.
@pytest.mark.parametrize("M,H", [(1, 768), (333, 768), (4192, 768), (100, 1024), (7, 256)])
def test_layernorm(M, H):
    g = torch.Generator(device="cpu").manual_seed(M + H)
    x = (torch.randn(M, H, generator=g) * 2 + 0.3).to(torch.bfloat16).to(DEV)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(H, generator=g)).to(DEV)
    out = ops.layernorm(x, gamma, beta, 1e-12)
    ref = torch.nn.functional.layer_norm(x.float(), (H,), gamma, beta, 1e-12)
    assert rel_err(out, ref) < 8e-3
    res = torch.randn(M, H, generator=g).to(torch.bfloat16).to(DEV)
    out = ops.layernorm(x, gamma, beta, 1e-12, resid=res)
    ref = torch.nn.functional.layer_norm(x.float() + res.float(), (H,), gamma, beta, 1e-12)
    assert rel_err(out, ref) < 8e-3


def test_timestep_mlp_matches_oracle():
    p = O.make_random_params(seed=1, seq_len=64)
    t = np.array([999.5, 3.0, 0.0, 500.0], dtype=np.float32)
    ref = O._linear(O._silu(O._linear(O.timestep_embedding(t, 128), p["time_embed.0.weight"], p["time_embed.0.bias"])),
                    p["time_embed.2.weight"], p["time_embed.2.bias"])
    dev = lambda k: torch.from_numpy(p[k]).to(DEV)
    out = ops.timestep_mlp(torch.from_numpy(t).to(DEV), dev("time_embed.0.weight"), dev("time_embed.0.bias"),
                           dev("time_embed.2.weight"), dev("time_embed.2.bias"))
    assert rel_err(out.cpu(), torch.from_numpy(ref)) < 1e-4


def test_embed_gather_and_cast():
    E = torch.randn(729, 128, device=DEV)
    for dt in (torch.int32, torch.int64):
        ids = torch.randint(0, 729, (3, 50), device=DEV, dtype=dt)
        assert torch.equal(ops.embed_gather(E, ids), E[ids.long()])
    x = torch.randn(5, 64, 128, device=DEV)
    assert torch.equal(ops.cast_bf16(x), x.to(torch.bfloat16))
    assert ops.embed_gather(E, torch.zeros((0, 4), dtype=torch.int64, device=DEV)).shape == (0, 4, 128)


# ------------------------------------------------------------------------------------------------ rounding
def test_rounding_golden(golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "rounding.npz"))
    E = torch.from_numpy(g["E"]).to(DEV)
    x = torch.from_numpy(g["x"]).to(DEV)
    idx, margin = ops.round_argmin(x, E, want_margin=True)
    idx = idx.cpu().numpy().astype(np.int64)
    ref_idx, dist = O.efficient_knn(g["E"], g["x"])
    ref_margin = O.top2_margin(dist)
    tol = 1e-3                                   # stated top-2 distance margin below which ids may differ
    differ = idx != g["idx"]
    assert not (differ & (ref_margin > tol)).any()
    assert idx[0] == 5 and idx[1] == 12          # duplicate rows -> lowest index
    assert np.array_equal(idx, g["idx"])         # in fact bit-exact on this fixture
    np.testing.assert_allclose(margin.cpu().numpy(), ref_margin, atol=2e-3)


@pytest.mark.parametrize("M", [1, 127, 128, 129, 5000])
def test_rounding_and_logits_random(M):
    rng = np.random.default_rng(M)
    E = rng.standard_normal((729, 128)).astype(np.float32)
    bias = rng.standard_normal(729).astype(np.float32)
    x = (rng.standard_normal((M, 128)) * 0.8).astype(np.float32)
    ref_idx, dist = O.efficient_knn(E, x)
    idx = ops.round_argmin(torch.from_numpy(x).to(DEV), torch.from_numpy(E).to(DEV)).cpu().numpy()
    bad = idx != ref_idx
    assert not (bad & (O.top2_margin(dist) > 1e-3)).any()
    logits = x @ E.T + bias
    tok, mg = ops.logits_argmax(torch.from_numpy(x).to(DEV), torch.from_numpy(E).to(DEV), torch.from_numpy(bias).to(DEV),
                                want_margin=True)
    tok = tok.cpu().numpy()
    srt = np.sort(logits, axis=1)
    bad = tok != logits.argmax(1)
    assert not (bad & ((srt[:, -1] - srt[:, -2]) > 1e-3)).any()
    assert ops.round_argmin(torch.zeros((0, 128), device=DEV), torch.from_numpy(E).to(DEV)).numel() == 0


@pytest.mark.parametrize("M", [1, 127, 128, 129, 5000, 70000])
def test_rounding_tensor_core(M, golden_dir):
    """tcgen05 split-bf16 rounding / decode: same ids as the fp32 oracle wherever the top-2 margin exceeds 1e-3."""
    rng = np.random.default_rng(M + 1)
    E = rng.standard_normal((729, 128)).astype(np.float32)
    E[700] = E[5]
    bias = rng.standard_normal(729).astype(np.float32)
    x = (rng.standard_normal((M, 128)) * 0.8).astype(np.float32)
    x[0] = E[700]                                      # exact hit on a duplicated row -> lowest index
    se = ops.SplitEmbedding(torch.from_numpy(E).to(DEV))
    idx, mg = ops.round_argmin_tc(torch.from_numpy(x).to(DEV), se, want_margin=True)
    idx = idx.cpu().numpy()
    ref_idx, dist = O.efficient_knn(E, x)
    ref_margin = O.top2_margin(dist)
    assert idx[0] == 5
    bad = idx != ref_idx
    assert not (bad & (ref_margin > 1e-3)).any(), (int(bad.sum()), float(ref_margin[bad].max()))
    np.testing.assert_allclose(mg.cpu().numpy()[~bad], ref_margin[~bad], atol=3e-3)
    tok = ops.round_argmin_tc(torch.from_numpy(x).to(DEV), se, cst=se.logit_cst(torch.from_numpy(bias).to(DEV)), mode=1).cpu().numpy()
    logits = x.astype(np.float64) @ E.T.astype(np.float64) + bias
    srt = np.sort(logits, axis=1)
    bad = tok != logits.argmax(1)
    assert not (bad & ((srt[:, -1] - srt[:, -2]) > 1e-3)).any()
    g = np.load(__import__("os").path.join(golden_dir, "rounding.npz"))
    seg = ops.SplitEmbedding(torch.from_numpy(g["E"]).to(DEV))
    assert np.array_equal(ops.round_argmin_tc(torch.from_numpy(g["x"]).to(DEV), seg).cpu().numpy(), g["idx"])


# ------------------------------------------------------------------------------------------------ posterior step
def _schedule(T=2000):
    s = O.make_schedule("sqrt", T)
    ops.set_schedule({n: getattr(s, n) for n in ops.TABLE_ORDER})
    return s


@pytest.mark.parametrize("tvals", [[1999, 1000, 1], [0, 0, 7]])
@pytest.mark.parametrize("mode", ["ddpm", "ddim"])
def test_posterior_step_matches_oracle(mode, tvals):
    s = _schedule()
    rng = np.random.default_rng(3)
    B, L, D = 3, 70, 128
    E = rng.standard_normal((729, D)).astype(np.float32)
    x = rng.standard_normal((B, L, D)).astype(np.float32)
    mo = rng.standard_normal((B, L, D)).astype(np.float32)
    noise = rng.standard_normal((B, L, D)).astype(np.float32)
    ids = rng.integers(0, 729, (B, L))
    x_start = E[ids]
    mask_tok = (rng.random((B, L)) > 0.3).astype(np.int64)
    mask = np.broadcast_to(mask_tok[..., None], (B, L, D))
    t = np.asarray(tvals, dtype=np.int64)
    if mode == "ddpm":
        ref = O.p_sample_step(s, x, t, mo, noise, E, True, mask, x_start)["sample"]
    else:
        ref = O.ddim_step(s, x, t, mo, noise, E, True, 0.0, mask, x_start)["sample"]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    idx = ops.round_argmin(dev(mo), dev(E))
    assert np.array_equal(idx.cpu().numpy(), O.efficient_knn(E, mo)[0])
    out_bf16 = torch.empty((B, L, D), dtype=torch.bfloat16, device=DEV)
    m = torch.broadcast_to(dev(mask_tok).unsqueeze(-1), (B, L, D))     # stride-0 expand, as run/sample.py:186 builds it
    out = ops.posterior_step(dev(x), dev(t), _lib.STEP_DDPM if mode == "ddpm" else _lib.STEP_DDIM, idx=idx, E=dev(E),
                             noise=dev(noise), mask=m, x_start=dev(x_start), clip=True, out_bf16=out_bf16)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-4, atol=1e-5)
    assert torch.equal(out_bf16, out.to(torch.bfloat16))
    # element-wise (non-broadcast) mask and the unrounded / unclipped variant
    mask_full = (rng.random((B, L, D)) > 0.5).astype(np.int64)
    if mode == "ddpm":
        ref2 = O.p_sample_step(s, x, t, mo, noise, None, False, mask_full, x_start)["sample"]
    else:
        ref2 = O.ddim_step(s, x, t, mo, noise, None, False, 0.5, mask_full, x_start)["sample"]
    out2 = ops.posterior_step(dev(x), dev(t), _lib.STEP_DDPM if mode == "ddpm" else _lib.STEP_DDIM, pred=dev(mo),
                              noise=dev(noise), mask=dev(mask_full), x_start=dev(x_start), clip=False,
                              eta=0.0 if mode == "ddpm" else 0.5)
    np.testing.assert_allclose(out2.cpu().numpy(), ref2, rtol=1e-4, atol=1e-5)


def test_q_sample_and_generation_init():
    s = _schedule()
    rng = np.random.default_rng(4)
    B, L, D = 2, 33, 128
    x0 = rng.standard_normal((B, L, D)).astype(np.float32)
    noise = rng.standard_normal((B, L, D)).astype(np.float32)
    mask_tok = (rng.random((B, L)) > 0.3).astype(np.int64)
    mask = np.broadcast_to(mask_tok[..., None], (B, L, D))
    t = np.array([74, 1999], dtype=np.int64)
    ref = O.q_sample(s, x0, t, noise, mask)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    out = ops.q_sample(dev(x0), dev(t), noise=dev(noise), mask=dev(mask_tok))
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-6, atol=1e-7)
    out = ops.q_sample(dev(x0), None, noise=dev(noise), mask=dev(mask_tok))
    assert np.array_equal(out.cpu().numpy(), np.where(mask == 0, x0, noise))


def test_xstart_from_eps():
    s = _schedule()
    rng = np.random.default_rng(6)
    x = rng.standard_normal((2, 9, 128)).astype(np.float32)
    e = rng.standard_normal((2, 9, 128)).astype(np.float32)
    t = np.array([5, 1500])
    ref = O.extract(s.sqrt_recip_alphas_cumprod, t, 3) * x - O.extract(s.sqrt_recipm1_alphas_cumprod, t, 3) * e
    out = ops.xstart_from_eps(torch.from_numpy(x).to(DEV), torch.from_numpy(e).to(DEV), torch.from_numpy(t).to(DEV))
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("top_p", [0.0, 1.0, 2.0])
def test_philox_noise_distribution(top_p):
    n = 1 << 22
    a = ops.fill_normal((n,), DEV, seed=105, step_counter=3, top_p=top_p)
    b = ops.fill_normal((n,), DEV, seed=105, step_counter=3, top_p=top_p)
    c = ops.fill_normal((n,), DEV, seed=105, step_counter=4, top_p=top_p)
    assert torch.equal(a, b) and not torch.equal(a, c)          # counter-based: reproducible, step-dependent
    # sharding invariance: the second half generated on its own equals the second half of the whole
    h = ops.fill_normal((n // 2,), DEV, seed=105, step_counter=3, elem_offset=n // 2, top_p=top_p)
    assert torch.equal(h, a[n // 2:])
    a = a.double().cpu().numpy()
    if top_p > 0:
        assert np.abs(a).max() <= top_p + 1e-6
        from scipy.stats import truncnorm
        var = truncnorm.var(-top_p, top_p)
    else:
        var = 1.0
        assert np.abs(a).max() > 4.0
    assert abs(a.mean()) < 3e-3 and abs(a.var() - var) < 5e-3
    assert abs(np.corrcoef(a[:-1], a[1:])[0, 1]) < 3e-3


def test_untruncated_normal_tails_are_symmetric_and_bounded():
    """2^28 untruncated draws hit both extreme 23-bit patterns with probability 1 - e^-32: the quantile is evaluated from the
    tail probability 0.5 - |q|, so both tails stop at |n| = Phi^-1(2^-24) = 5.30 (a quantile taken from p = q + 0.5 rounds the
    topmost draw to p = 1 and yields a +11.5 sigma outlier)."""
    n = 1 << 28
    a = ops.fill_normal((n,), DEV, seed=7, step_counter=1, top_p=0.0)
    hi, lo = float(a.max()), float(a.min())
    assert 5.2 < hi < 5.4 and -5.4 < lo < -5.2, (lo, hi)
    assert abs(hi + lo) < 0.05, (lo, hi)
    for thr in (3.0, 4.0):                                       # tail masses agree with the normal law on both sides
        up, dn = int((a > thr).sum()), int((a < -thr).sum())
        want = n * 0.5 * math.erfc(thr / math.sqrt(2.0))
        assert abs(up - want) < 6 * math.sqrt(want) and abs(dn - want) < 6 * math.sqrt(want), (thr, up, dn, want)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices in one process")
def test_one_process_two_devices():
    """The library keeps its state (schedule tables, SM count, raised shared-memory attributes) per device: one process that
    alternates between cuda:0 and cuda:1 gets the right tables and launchable tcgen05 kernels on both."""
    from musediffusion_b200.diffusion import SpacedDiffusion, get_named_beta_schedule, space_timesteps
    outs = []
    for dev_i, (sched, T) in enumerate([("sqrt", 2000), ("linear", 500)]):
        dev = torch.device("cuda", dev_i)
        with torch.cuda.device(dev):
            d = SpacedDiffusion(use_timesteps=space_timesteps(T, [T]), betas=get_named_beta_schedule(sched, T), rescale_timesteps=True,
                                predict_xstart=True)
            g = torch.Generator(device="cpu").manual_seed(dev_i)
            x0 = torch.randn(2, 40, 128, generator=g).to(dev)
            n = torch.randn(2, 40, 128, generator=g).to(dev)
            t = torch.tensor([T - 1, 3], device=dev)
            q = d.q_sample(x0, t, noise=n)
            s = O.make_schedule(sched, T)
            want = O.q_sample(s, x0.cpu().numpy(), t.cpu().numpy(), n.cpu().numpy())
            np.testing.assert_allclose(q.cpu().numpy(), want, rtol=1e-6, atol=1e-7)
            outs.append((d, x0, t, n, want, dev))
    # second round in the opposite order, no re-upload: device 0's tables must not have been displaced by device 1's
    for d, x0, t, n, want, dev in outs[::-1] + outs:
        with torch.cuda.device(dev):
            np.testing.assert_allclose(d.q_sample(x0, t, noise=n).cpu().numpy(), want, rtol=1e-6, atol=1e-7)
    for dev_i in (1, 0, 1):
        dev = torch.device("cuda", dev_i)
        with torch.cuda.device(dev):
            A = (torch.randn(300, 768, device=dev) * 0.1).to(torch.bfloat16)
            W = (torch.randn(768, 768, device=dev) * 0.1).to(torch.bfloat16)
            y = ops.linear(A, W, None)
            assert rel_err(y, A.float() @ W.float().T) < 1.2e-2
            qkv = (torch.randn(2 * 200, 3 * 128, device=dev) * 0.5).to(torch.bfloat16)
            o = ops.attention(qkv, 2, 200, 2)
            q_, k_, v_ = [t_.float().view(2, 200, 2, 64).transpose(1, 2) for t_ in qkv.split(128, dim=1)]
            ref = (torch.softmax(q_ @ k_.transpose(-1, -2), dim=-1) @ v_).transpose(1, 2).reshape(400, 128)
            assert rel_err(o, ref) < 1.5e-2
            x = torch.randn(500, 128, device=dev)
            E = torch.randn(729, 128, device=dev)
            assert torch.equal(ops.round_argmin_tc(x, ops.SplitEmbedding(E)), ops.round_argmin(x, E))


@pytest.mark.parametrize("M", [5000, 128, 77])
def test_linear_split_epilogue_equals_split_of_fp32_output(M):
    """MD_EPI_BIAS_SPLIT (last Linear of output_down_proj, network.py:85): bf16 [M, 2N] = [hi | lo] written by the GEMM epilogue
    is bit-identical to md_split_bf16 of the fp32 output, and rounds to the same ids without the split pass."""
    g = torch.Generator(device="cpu").manual_seed(M)
    A = (torch.randn(M, 768, generator=g) * 0.3).to(torch.bfloat16).to(DEV)
    W = (torch.randn(128, 768, generator=g) * 0.05).to(torch.bfloat16).to(DEV)
    b = torch.randn(128, generator=g).to(DEV)
    y32 = ops.linear(A, W, b, _lib.EPI_BIAS, out_dtype=torch.float32)
    want = ops.split_bf16(y32, copies=1)
    got = ops.linear(A, W, b, _lib.EPI_BIAS_SPLIT)
    assert got.shape == (M, 256) and got.dtype == torch.bfloat16
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    E = torch.randn(729, 128, generator=g).to(DEV)
    se = ops.SplitEmbedding(E)
    assert torch.equal(ops.round_argmin_tc(None, se, presplit=got), ops.round_argmin_tc(y32, se))
    with pytest.raises(_lib.MuseDiffLibraryError):
        ops.linear(A, W, b, _lib.EPI_BIAS_SPLIT, out=torch.empty(M, 128, dtype=torch.bfloat16, device=DEV))
