"""CPU tests: the numpy oracle (oracle/musediff_oracle.py) against fixtures produced by the UNMODIFIED
reference (oracle/make_golden.py) and against the known-answer values of SURVEY.md Appendix B."""
import os

import numpy as np
import pytest

import musediff_oracle as O


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


TABLES = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
          "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
          "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]


@pytest.mark.parametrize("sched,T,resp", [("sqrt", 2000, ""), ("linear", 1000, ""), ("cosine", 500, ""),
                                          ("trunc_cos", 400, ""), ("trunc_lin", 300, ""), ("pw_lin", 200, ""),
                                          ("sqrt", 2000, "ddim50"), ("sqrt", 300, "10,15,20")])
def test_schedule_tables_bit_exact(golden_dir, sched, T, resp):
    g = load(golden_dir, "schedules.npz")
    key = "%s_%d_%s" % (sched, T, resp.replace(",", "-") or "full")
    s = O.make_schedule(sched, T, resp)
    assert np.array_equal(O.get_named_beta_schedule(sched, T), g[key + "/raw_betas"])
    assert np.array_equal(np.asarray(s.timestep_map), g[key + "/timestep_map"])
    for n in TABLES:
        assert np.array_equal(getattr(s, n), g[key + "/" + n]), n


def test_schedule_known_answers():
    """SURVEY.md Appendix B."""
    s = O.make_schedule("sqrt", 2000)
    assert s.betas[0] == 0.01464131053316331
    assert s.betas[1] == 0.008889087768847226
    assert s.betas[1999] == 0.999
    assert s.alphas_cumprod[0] == 0.9853586894668367
    assert s.alphas_cumprod[1000] == 0.29542331509863634
    assert s.alphas_cumprod[1999] == 2.0204040808180702e-07
    assert s.posterior_mean_coef1[0] == 1.0 and s.posterior_mean_coef2[0] == 0.0
    assert s.posterior_mean_coef1[1] == 0.37708031820760624
    assert s.posterior_mean_coef2[1] == 0.6229032199965542
    assert s.posterior_mean_coef1[1999] == 0.014199880660890406
    assert s.posterior_mean_coef2[1999] == 0.03161639391078364
    assert s.model_variance[0] == 0.005561816310213409 and s.model_variance[1999] == 0.999
    assert O.space_timesteps(2000, [2000]) == set(range(2000))
    idx = O.ddim_indices(s, gap=20, t_enc=100)
    assert idx[0] == 1999 and idx[1] == 1979 and idx[-1] == 19 and len(idx) == 100
    assert float(s.model_timestep(np.array([1999]))[0]) == 999.5


def test_rounding_matches_reference(golden_dir):
    g = load(golden_dir, "rounding.npz")
    idx, dist = O.efficient_knn(g["E"], g["x"])
    assert np.array_equal(idx, g["idx"])
    assert idx[0] == 5 and idx[1] == 12            # duplicate rows 700/5 and 300/12 -> lowest index
    np.testing.assert_allclose(-dist.min(0), g["val"], rtol=0, atol=2e-4)
    assert np.array_equal(O.denoised_fn_round(g["E"], g["x"]), g["rounded"])


def test_meta_prefix_format(golden_dir):
    g = load(golden_dir, "meta_batch.npz")
    prefix = [574, 623, 627, 635, 639, 642, 651, 684, 694, 720, 727, 432, 199, 432, 285, 432, 267, 432, 258,
              432, 199, 432, 285, 432, 267, 432, 258]
    assert g["input_ids"].dtype == np.int32 and g["input_ids"][0, :27].tolist() == prefix
    assert (g["input_mask"][:, :28] == 0).all() and (g["input_mask"][:, 28:] == 1).all()
    b = O.make_synthetic_batch("generation", 2, 64, seed=1)
    assert b["input_ids"].dtype == np.int32 and b["input_mask"].dtype == np.int32
    n = int((b["input_mask"][0] == 0).sum())
    assert (b["input_ids"][0, n - 1:] == 0).all() and (b["input_ids"][0, :n - 1] > 0).all()


@pytest.mark.parametrize("name", ["forward_tiny.npz", "forward_ragged.npz"])
def test_denoiser_forward_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    p = O.make_random_params(seed=int(g["seed"]), seq_len=int(g["seq_len"]))
    out = O.denoiser_forward(p, g["x"], g["t"])
    assert rel(out, g["model_output"]) < 2e-5, rel(out, g["model_output"])
    emb_t = O._linear(O._silu(O._linear(O.timestep_embedding(g["t"], 128), p["time_embed.0.weight"],
                                        p["time_embed.0.bias"])), p["time_embed.2.weight"], p["time_embed.2.bias"])
    assert rel(emb_t, g["emb_t"]) < 1e-5
    assert np.array_equal(O.logits_argmax(p, g["x"]), g["tokens"])


def test_denoiser_forward_base_shape(golden_dir):
    g = load(golden_dir, "forward_base.npz")
    p = O.make_random_params(seed=int(g["seed"]), seq_len=int(g["seq_len"]))
    out = O.denoiser_forward(p, g["x"], g["t"])
    assert out.shape == (1, 2096, 128)
    assert rel(out, g["model_output"]) < 5e-5, rel(out, g["model_output"])


def test_single_steps_match_reference(golden_dir):
    g = load(golden_dir, "steps_tiny.npz")
    seed, L = int(g["seed"]), int(g["seq_len"])
    p = O.make_random_params(seed=seed, seq_len=L)
    s = O.make_schedule()
    E = p["word_embedding.weight"]
    x_start = O.get_embeds(p, g["input_ids"])
    mask = np.broadcast_to(g["input_mask"][..., None], x_start.shape)
    B = x_start.shape[0]
    for tag, tval in [("t1999", 1999), ("t1000", 1000), ("t1", 1), ("t0", 0)]:
        x = g[tag + "/x"]
        t = np.full((B,), tval, dtype=np.int64)
        mo = g[tag + "/model_output"]
        n = O.NoiseStream(seed * 1000 + tval).truncated(x.shape, 1)
        r = O.p_sample_step(s, x, t, mo, n, E, True, mask, x_start)
        np.testing.assert_allclose(r["sample"], g[tag + "/p_sample"], rtol=1e-5, atol=1e-6)
        assert np.array_equal(r["pred_xstart"], g[tag + "/pred_xstart"])
        n2 = O.NoiseStream(seed * 1000 + tval + 1).randn(x.shape)
        r2 = O.ddim_step(s, x, t, mo, n2, E, True, 0.0, mask, x_start)
        np.testing.assert_allclose(r2["sample"], g[tag + "/ddim_sample"], rtol=1e-4, atol=2e-5)
        n3 = O.NoiseStream(seed * 1000 + tval + 2).truncated(x.shape, 0)
        r3 = O.p_sample_step(s, x, t, mo, n3, None, False, None, None)
        np.testing.assert_allclose(r3["sample"], g[tag + "/p_sample_raw"], rtol=1e-5, atol=1e-6)
        # the oracle's own denoiser reproduces the model output that the reference produced
        mo2 = O.denoiser_forward(p, x, s.model_timestep(t))
        assert rel(mo2, mo) < 5e-5
    n = O.NoiseStream(seed * 1000 + 77).randn(x_start.shape + (1,))
    xq = O.q_sample(s, x_start[..., None], np.full((B, 1), 74), n, mask[..., None])[..., 0]
    np.testing.assert_allclose(xq, g["q_sample_t74"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["loop_gen_ddpm.npz", "loop_mod_ddpm.npz", "loop_mod_ddim.npz", "loop_gen_ddim.npz"])
def test_full_loops_match_reference(golden_dir, name):
    """whole run/sample.py:177-220 slice: identical noise stream -> identical decoded tokens."""
    g = load(golden_dir, name)
    seed, L = int(g["seed"]), int(g["seq_len"])
    p = O.make_random_params(seed=seed, seq_len=L)
    s = O.make_schedule("sqrt", int(g["diffusion_steps"]))
    cond = {"input_ids": g["input_ids"], "input_mask": g["input_mask"]}
    rec = []
    tokens = O.sample_batch(s, p, cond, str(g["mode"]), int(g["step"]), O.NoiseStream(seed + 999),
                            strength=float(g["strength"]), top_p=int(g["top_p"]), record=rec)
    np.testing.assert_allclose(rec[0]["x_t"], g["x_noised"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rec[-1]["sample"], g["final_sample"], rtol=1e-4, atol=1e-5)
    assert np.array_equal(tokens, g["tokens"])
    n_masked = (g["input_mask"] == 0)
    assert np.array_equal(tokens[n_masked][g["input_ids"][n_masked] > 0],
                          g["input_ids"][n_masked][g["input_ids"][n_masked] > 0])
