"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Drives the UNMODIFIED reference (`/root/reference` in the build container, the git-ignored copy `oracle/_ref/` made
by `oracle/build_ref.sh` on the GPU box) through its own entry point `MuseDiffusion.run.sample.main()`
(run/sample.py:22-311) and through the driver slice run/sample.py:177-220, in two arms:

  * reference arm — stock modules, forced onto the host CPU (the reference's fp32 torch path);
  * drop-in arm   — the three `sys.modules` swaps of INTEGRATION.md, everything else of `main()` untouched, on cuda:0.

Both arms consume the same numpy noise stream (`torch.randn_like` is patched for the reference's own calls,
`GaussianDiffusion.noise_source` feeds the CUDA sampler), both record ids + top-2 margins of every rounding call, and
the MIDI tail (`decode_batch`, needs miditoolkit) is replaced by a recorder of the token batches, as SURVEY.md
Appendix A prescribes.  Used by tests/test_dropin_gpu.py and by bench.py's reference / cpu_baseline legs only.
"""
import contextlib
import os
import sys
import types
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import musediff_oracle as O  # noqa: E402
import ref_shim  # noqa: E402

SWAPPED = ("MuseDiffusion.models.diffusion", "MuseDiffusion.models.network", "MuseDiffusion.models.rounding")


def available():
    return ref_shim.reference_available()


def install():
    ref_shim.install_reference_shim()


def write_checkpoint_dir(dirpath, params, seq_len, diffusion_steps=2000, **overrides):
    """A model directory as the reference's trainer leaves it: `model_000000.pt` (state dict, utils/train_util.py) and
    `training_args.json` written by the reference's own `TrainSettings(...).json()` (config/train.py:101)."""
    import torch
    from MuseDiffusion.config import TrainSettings
    os.makedirs(dirpath, exist_ok=True)
    state = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params.items()}
    model_path = os.path.join(dirpath, "model_000000.pt")
    torch.save(state, model_path)
    settings = TrainSettings(seq_len=seq_len, diffusion_steps=diffusion_steps, use_corruption=False, **overrides)
    with open(os.path.join(dirpath, "training_args.json"), "w") as f:
        f.write(settings.json())
    return model_path


class _patched_randn_like:
    """torch.randn_like -> NoiseStream draws (on the input's device), for the reference's own calls."""

    def __init__(self, stream):
        self.stream = stream

    def __enter__(self):
        import torch
        self.orig = torch.randn_like
        torch.randn_like = lambda x, **kw: torch.from_numpy(self.stream.randn(tuple(x.shape))).to(x.device)
        return self

    def __exit__(self, *a):
        import torch
        torch.randn_like = self.orig


@contextlib.contextmanager
def _dropin_swaps():
    """re-point the three modules the sampling path resolves; nothing else of the reference changes."""
    import musediffusion_b200.diffusion as d
    import musediffusion_b200.network as n
    import musediffusion_b200.rounding as r
    saved = {k: sys.modules.get(k) for k in SWAPPED}
    sys.modules[SWAPPED[0]], sys.modules[SWAPPED[1]], sys.modules[SWAPPED[2]] = d, n, r
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def run_main(mode, model_path, out_dir, extra_argv=(), batches=None, dropin=False, stream_seed=0, top_p=1, midi_tail=False):
    """MuseDiffusion.run.sample.main(namespace), unmodified.  Returns dict(tokens=[...], masks=[...], step_ids, step_margin).
    `midi_tail` (modification mode only — generation loops until enough rows are valid): `decode_batch` resolves to the
    package's instead of the recorder, so the run ends in MIDI files.
    `batches`: modification mode's data loader replacement (list of cond dicts), SURVEY.md §8c item 4."""
    import torch
    install()
    import MuseDiffusion.run.sample as ref_sample
    import MuseDiffusion.utils.decode_util as ref_decode
    import MuseDiffusion.utils.dist_util as ref_dist
    import MuseDiffusion.data as ref_data
    import MuseDiffusion.models.rounding as ref_rounding

    captured = {"tokens": [], "masks": []}

    def fake_decode_batch(mode, sequences, input_ids_mask_ori, output_dir, batch_index, previous_count, **kw):
        captured["tokens"].append(np.array(sequences))
        captured["masks"].append(np.array(input_ids_mask_ori))
        if midi_tail:                                             # the fourth swap: the package's decode_batch writes the files
            import musediffusion_b200.decode_util as our_decode
            out = our_decode.decode_batch(mode, sequences, input_ids_mask_ori, batch_index, previous_count, output_dir,
                                          return_indices=True, strict_validation=kw.get("strict_validation", False))
            captured.setdefault("valid", []).append(out)
            captured["output_dir"] = output_dir
            return out
        return len(sequences), []

    stream = O.NoiseStream(stream_seed)
    rec_ids, rec_margin = [], []
    argv = [mode, "--model_path", model_path, "--out_dir", out_dir] + list(extra_argv)
    ns = ref_sample.create_parser().parse_args(argv)
    saved = (ref_decode.decode_batch, ref_data.load_data_music, ref_dist._cuda_available, ref_rounding.denoised_fn_round)
    ref_decode.decode_batch = fake_decode_batch
    if batches is not None:
        ref_data.load_data_music = lambda **kw: list(batches)
    ref_dist.setup_dist.cache_clear()
    try:
        if dropin:
            import musediffusion_b200.diffusion as ours
            ours.GaussianDiffusion.noise_source = staticmethod(
                lambda shape, kind: stream.truncated(shape, top_p) if kind == "truncated" else stream.randn(shape))
            trace = []
            ours.GaussianDiffusion.rounding_trace = trace
            try:
                with _dropin_swaps(), _patched_randn_like(stream):
                    ref_sample.main(ns)
            finally:
                ours.GaussianDiffusion.noise_source = None
                ours.GaussianDiffusion.rounding_trace = None
            B, L = captured["tokens"][0].shape
            for ids, margin in trace:
                rec_ids.append(ids.view(-1, L).cpu().numpy())
                rec_margin.append(margin.view(-1, L).cpu().numpy())
        else:
            ref_dist._cuda_available = lambda: False              # reference arm = the reference's CPU path
            inner = ref_rounding.denoised_fn_round

            def recording_round(model_emb, text_emb, t, dist=None):
                E = model_emb.weight
                flat = text_emb.reshape(-1, text_emb.size(-1))
                d = torch.clamp((E ** 2).sum(-1).view(-1, 1) + (flat ** 2).sum(-1).view(1, -1)
                                - 2.0 * torch.mm(E, flat.transpose(0, 1)), 0.0, np.inf)
                two = torch.topk(-d, k=2, dim=0)
                rec_margin.append((two.values[0] - two.values[1]).cpu().numpy().reshape(text_emb.shape[:-1]))
                rec_ids.append(ref_rounding.get_efficient_knn(E, flat)[1][0].cpu().numpy().reshape(text_emb.shape[:-1]))
                return inner(model_emb, text_emb, t, dist=dist)

            ref_rounding.denoised_fn_round = recording_round
            with _patched_randn_like(stream):
                ref_sample.main(ns)
    finally:
        ref_decode.decode_batch, ref_data.load_data_music, ref_dist._cuda_available, ref_rounding.denoised_fn_round = saved
        ref_dist.setup_dist.cache_clear()
    captured["step_ids"] = rec_ids
    captured["step_margin"] = rec_margin
    return captured


# ------------------------------------------------------------------------------------------------ CPU baseline
class ReferenceSampler:
    """The reference's sampling slice (run/sample.py:84-114 set-up, :177-220 per batch) on the host CPU with the
    reference's own modules — what BASELINE.md section 3 times.  No module of musediffusion_b200 is involved."""

    def __init__(self, seq_len=2096, diffusion_steps=2000, seed=0, threads=None):
        import torch
        install()
        from MuseDiffusion.utils.initialization import create_model_and_diffusion
        torch.set_num_threads(threads or os.cpu_count())
        self.threads = torch.get_num_threads()
        args = types.SimpleNamespace(hidden_dim=128, hidden_t_dim=128, vocab_size=729, seq_len=seq_len, dropout=0.1,
                                     noise_schedule="sqrt", diffusion_steps=diffusion_steps, timestep_respacing="",
                                     rescale_timesteps=True, predict_xstart=True)
        torch.manual_seed(seed)
        self.model, self.diffusion = create_model_and_diffusion(args)          # random init, as BASELINE.md section 3
        self.model.eval().requires_grad_(False)
        self.model_emb = torch.nn.Embedding(num_embeddings=729, embedding_dim=128, padding_idx=0,
                                            _weight=self.model.word_embedding.weight.clone().cpu())
        self.model_emb.eval().requires_grad_(False)
        self.seq_len, self.T = seq_len, diffusion_steps

    def sample(self, cond, mode, step, strength=1.0, top_p=1, n_steps=None):
        """run/sample.py:177-220; `n_steps` truncates the chain through the reference's own `t_enc` argument (a bounded
        sample of the chain: the first n_steps indices from the top, every step costs the same)."""
        import torch
        from MuseDiffusion.models.rounding import denoised_fn_round
        model, diffusion = self.model, self.diffusion
        with torch.no_grad():
            ids = torch.as_tensor(cond["input_ids"])
            mask_ori = torch.as_tensor(cond["input_mask"])
            x_start = model.get_embeds(ids)
            input_ids_mask = torch.broadcast_to(mask_ori.unsqueeze(dim=-1), x_start.shape)
            if mode == "generation":
                noising_t = None
                x_noised = torch.where(torch.eq(input_ids_mask, 0), x_start, torch.randn_like(x_start))
            else:
                noising_t = int(step * strength)
                timestep = torch.full((len(ids), 1), noising_t - 1)
                x_noised = diffusion.q_sample(x_start.unsqueeze(-1), timestep, mask=input_ids_mask).squeeze(-1)
            if n_steps is not None:
                noising_t = n_steps if noising_t is None else min(noising_t, n_steps)
            if step == self.T:
                gap, sample_fn = 1, diffusion.p_sample_loop
            else:
                gap, sample_fn = self.T // step, diffusion.ddim_sample_loop
            samples = sample_fn(model=model, shape=tuple(x_start.shape), noise=x_noised, clip_denoised=True,
                                denoised_fn=partial(denoised_fn_round, self.model_emb, dist=None), model_kwargs=cond,
                                top_p=top_p, clamp_step=0, clamp_first=True, mask=input_ids_mask, x_start=x_start,
                                gap=gap, t_enc=noising_t, only_last=True)
            return torch.argmax(model.get_logits(samples[-1]), dim=-1)
