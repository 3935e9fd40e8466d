"""
ORACLE — TEST INFRASTRUCTURE ONLY (build-container only).

Generates `tests/golden/*.npz` by running the UNMODIFIED reference (`/root/reference`) through
`oracle/ref_shim.py`.  Run from the repo root:   python oracle/make_golden.py

Weights come from `musediff_oracle.make_random_params(seed, ...)` (numpy PCG64, so the tests can
re-create them without torch RNG), loaded into the reference module with `load_state_dict`.
Noise: `torch.randn_like` is monkey-patched for the duration of a reference call so that it draws
from `musediff_oracle.NoiseStream(seed)`; the oracle mirrors the reference's call order, so both
consume identical noise and whole loops become comparable.
"""
import os
import sys
import types
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import musediff_oracle as O  # noqa: E402
from ref_shim import install_reference_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def to_torch_state(p):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in p.items()}


def build_reference(p, seq_len, diffusion_steps=2000, noise_schedule="sqrt", timestep_respacing="",
                    rescale_timesteps=True, predict_xstart=True):
    import torch
    from MuseDiffusion.utils.initialization import create_model_and_diffusion
    args = types.SimpleNamespace(hidden_dim=128, hidden_t_dim=128, vocab_size=729, seq_len=seq_len, dropout=0.1,
                                 noise_schedule=noise_schedule, diffusion_steps=diffusion_steps,
                                 timestep_respacing=timestep_respacing, rescale_timesteps=rescale_timesteps,
                                 predict_xstart=predict_xstart)
    model, diffusion = create_model_and_diffusion(args)
    missing = model.load_state_dict(to_torch_state(p), strict=True)
    model.eval().requires_grad_(False)
    model_emb = torch.nn.Embedding(num_embeddings=729, embedding_dim=128, padding_idx=0,
                                   _weight=model.word_embedding.weight.clone().cpu())
    model_emb.eval().requires_grad_(False)
    return model, diffusion, model_emb


class patched_randn_like:
    def __init__(self, stream):
        self.stream = stream

    def __enter__(self):
        import torch
        self.orig = torch.randn_like
        torch.randn_like = lambda x, **kw: torch.from_numpy(self.stream.randn(tuple(x.shape)))
        return self

    def __exit__(self, *a):
        import torch
        torch.randn_like = self.orig


def golden_schedules():
    from MuseDiffusion.models import diffusion as D
    out = {}
    names = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
             "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
             "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]
    for sched, T, resp in [("sqrt", 2000, ""), ("linear", 1000, ""), ("cosine", 500, ""), ("trunc_cos", 400, ""),
                           ("trunc_lin", 300, ""), ("pw_lin", 200, ""), ("sqrt", 2000, "ddim50"),
                           ("sqrt", 300, "10,15,20")]:
        betas = D.get_named_beta_schedule(sched, T)
        use = D.space_timesteps(T, resp if resp else [T])
        d = D.SpacedDiffusion(use_timesteps=use, betas=betas, rescale_timesteps=True, predict_xstart=True)
        key = "%s_%d_%s" % (sched, T, resp.replace(",", "-") or "full")
        out[key + "/raw_betas"] = betas
        out[key + "/timestep_map"] = np.asarray(d.timestep_map, dtype=np.int64)
        for n in names:
            out[key + "/" + n] = np.asarray(getattr(d, n), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **out)
    print("schedules.npz", len(out), "arrays")


def golden_rounding():
    import torch
    from MuseDiffusion.models.rounding import get_efficient_knn, denoised_fn_round
    rng = np.random.default_rng(7)
    E = rng.standard_normal((729, 128)).astype(np.float32)
    E[700] = E[5]                                           # duplicate rows -> tie -> lowest index
    E[300] = E[12]
    x = rng.standard_normal((6, 37, 128)).astype(np.float32) * 0.7
    x[0, 0] = E[700]                                        # exact hits (distance ~0 before clamp)
    x[0, 1] = E[300] + 1e-4
    x[1, 2] = 0.0
    x[2, :5] = E[[0, 728, 1, 5, 12]]
    val, idx = get_efficient_knn(torch.from_numpy(E), torch.from_numpy(x.reshape(-1, 128)))
    emb = torch.nn.Embedding(729, 128, _weight=torch.from_numpy(E.copy()))
    new = denoised_fn_round(emb, torch.from_numpy(x), None)
    np.savez_compressed(os.path.join(OUT, "rounding.npz"), E=E, x=x, idx=idx[0].numpy(), val=val[0].numpy(),
                        rounded=new.detach().numpy())
    print("rounding.npz")


def golden_forward(seq_len, B, seed, fname, tvals):
    import torch
    p = O.make_random_params(seed=seed, seq_len=seq_len)
    model, diffusion, _ = build_reference(p, seq_len)
    rng = np.random.default_rng(seed + 100)
    x = rng.standard_normal((B, seq_len, 128)).astype(np.float32)
    t = np.asarray(tvals, dtype=np.float32)
    with torch.no_grad():
        out = model(torch.from_numpy(x), torch.from_numpy(t)).numpy()
        emb_t = model.time_embed(model.timestep_embedding(torch.from_numpy(t), 128)).numpy()
        logits = model.get_logits(torch.from_numpy(x)).numpy()
    np.savez_compressed(os.path.join(OUT, fname), seed=seed, seq_len=seq_len, x=x, t=t, model_output=out,
                        emb_t=emb_t, tokens=logits.argmax(-1))
    print(fname, out.shape, float(np.abs(out).mean()))


def golden_steps(seq_len=64, B=3, seed=3):
    """single p_sample / ddim_sample / q_sample calls on the reference with injected noise."""
    import torch
    p = O.make_random_params(seed=seed, seq_len=seq_len)
    model, diffusion, model_emb = build_reference(p, seq_len)
    rng = np.random.default_rng(seed + 200)
    cond = O.make_synthetic_batch("modification", B, seq_len, seed=seed)
    ids = torch.from_numpy(cond["input_ids"])
    x_start = model.get_embeds(ids)
    mask = torch.broadcast_to(torch.from_numpy(cond["input_mask"]).unsqueeze(-1), x_start.shape)
    out = {"seed": seed, "seq_len": seq_len, "input_ids": cond["input_ids"], "input_mask": cond["input_mask"]}
    for tag, tval in [("t1999", 1999), ("t1000", 1000), ("t1", 1), ("t0", 0)]:
        x = torch.from_numpy(rng.standard_normal((B, seq_len, 128)).astype(np.float32))
        x = torch.where(mask == 0, x_start, x)
        t = torch.tensor([tval] * B)
        stream = O.NoiseStream(seed * 1000 + tval)
        with patched_randn_like(stream), torch.no_grad():
            r = diffusion.p_sample(model, x, t, clip_denoised=True,
                                   denoised_fn=partial(__import__("MuseDiffusion.models.rounding", fromlist=["x"]).denoised_fn_round, model_emb, dist=None),
                                   model_kwargs={}, top_p=1, mask=mask, x_start=x_start)
        out[tag + "/x"] = x.numpy()
        out[tag + "/p_sample"] = r["sample"].numpy()
        out[tag + "/pred_xstart"] = r["pred_xstart"].numpy()
        out[tag + "/model_output"] = diffusion._wrap_model(model)(x, t).numpy()
        stream = O.NoiseStream(seed * 1000 + tval + 1)
        with patched_randn_like(stream), torch.no_grad():
            r2 = diffusion.ddim_sample(model, x, t, clip_denoised=True,
                                       denoised_fn=partial(__import__("MuseDiffusion.models.rounding", fromlist=["x"]).denoised_fn_round, model_emb, dist=None),
                                       model_kwargs={}, mask=mask, x_start=x_start, eta=0.0)
        out[tag + "/ddim_sample"] = r2["sample"].numpy()
        # unclamped / unrounded variants (denoised_fn=None, clip_denoised False) with top_p=0
        stream = O.NoiseStream(seed * 1000 + tval + 2)
        with patched_randn_like(stream), torch.no_grad():
            r3 = diffusion.p_sample(model, x, t, clip_denoised=False, denoised_fn=None, model_kwargs={},
                                    top_p=0, mask=None, x_start=None)
        out[tag + "/p_sample_raw"] = r3["sample"].numpy()
    # q_sample as run/sample.py:195-197 calls it
    stream = O.NoiseStream(seed * 1000 + 77)
    with patched_randn_like(stream):
        tq = torch.full((B, 1), 74)
        xq = diffusion.q_sample(x_start.unsqueeze(-1), tq, mask=mask).squeeze(-1)
    out["q_sample_t74"] = xq.numpy()
    np.savez_compressed(os.path.join(OUT, "steps_tiny.npz"), **out)
    print("steps_tiny.npz")


def golden_loop(fname, mode, seq_len, B, seed, diffusion_steps, step, strength=0.75, top_p=1):
    """run/sample.py:177-220 replayed with reference objects (the data loader / MIDI tail are out of scope)."""
    golden_loop_base(fname, mode, B, seed, diffusion_steps, step, strength=strength, top_p=top_p, seq_len=seq_len,
                     store_x_noised=True)


def golden_loop_base(fname, mode, B, seed, diffusion_steps, step, strength=1.0, top_p=1, seq_len=2096,
                     store_x_noised=None, store_final=True):
    """BASELINE.json configs[0] / configs[2] at the base sequence length: the driver slice run/sample.py:177-220 on
    the unmodified reference, recording for EVERY rounding call (rounding.py:31-47, invoked at diffusion.py:322)
    the chosen ids and the top-2 squared-distance margin, so the GPU parity test can apply the north-star rule
    "tokens bit-exact except where the top-2 margin is below tolerance" over the whole chain.
    x_noised is stored for modification mode (q_sample arithmetic); in generation mode it is a pure select of the
    NoiseStream's first draw and the test regenerates it."""
    import time
    import torch
    from MuseDiffusion.models.rounding import denoised_fn_round, get_efficient_knn
    p = O.make_random_params(seed=seed, seq_len=seq_len)
    model, diffusion, model_emb = build_reference(p, seq_len, diffusion_steps=diffusion_steps)
    cond = O.make_synthetic_batch(mode, B, seq_len, seed=seed + 5)
    ids = torch.from_numpy(cond["input_ids"])
    mask_ori = torch.from_numpy(cond["input_mask"])
    stream = O.NoiseStream(seed + 999)
    rec_ids, rec_margin = [], []
    inner = partial(denoised_fn_round, model_emb, dist=None)

    def recording_round(x, t):
        E = model_emb.weight
        flat = x.reshape(-1, x.size(-1))
        emb_norm = (E ** 2).sum(-1).view(-1, 1)
        arr_norm = (flat ** 2).sum(-1).view(-1, 1)
        dist = torch.clamp(emb_norm + arr_norm.transpose(0, 1) - 2.0 * torch.mm(E, flat.transpose(0, 1)), 0.0, np.inf)
        two = torch.topk(-dist, k=2, dim=0)
        rec_margin.append((two.values[0] - two.values[1]).numpy().astype(np.float32).reshape(x.shape[:-1]))
        _, idx = get_efficient_knn(E, flat)
        rec_ids.append(idx[0].numpy().astype(np.int16).reshape(x.shape[:-1]))
        return inner(x, t)

    if step == diffusion_steps:
        gap, sample_fn = 1, diffusion.p_sample_loop
    else:
        gap, sample_fn = diffusion_steps // step, diffusion.ddim_sample_loop
    t0 = time.time()
    with patched_randn_like(stream), torch.no_grad():
        x_start = model.get_embeds(ids)
        input_ids_mask = torch.broadcast_to(mask_ori.unsqueeze(dim=-1), x_start.shape)
        if mode == "generation":
            noising_t = None
            noise = torch.randn_like(x_start)
            x_noised = torch.where(torch.eq(input_ids_mask, 0), x_start, noise)
        else:
            noising_t = int(step * strength)
            timestep = torch.full((B, 1), noising_t - 1)
            x_noised = diffusion.q_sample(x_start.unsqueeze(-1), timestep, mask=input_ids_mask).squeeze(-1)
        samples = sample_fn(model=model, shape=(B, seq_len, 128), noise=x_noised, clip_denoised=True,
                            denoised_fn=recording_round, model_kwargs=cond,
                            top_p=top_p, clamp_step=0, clamp_first=True, mask=input_ids_mask, x_start=x_start,
                            gap=gap, t_enc=noising_t, only_last=True)
        sample = samples[-1]
        logits = model.get_logits(sample)
        tokens = torch.argmax(logits, dim=-1)
        top2 = torch.topk(logits, k=2, dim=-1).values
    if store_x_noised is None:
        store_x_noised = mode != "generation"
    extra = {"x_noised": x_noised.numpy()} if store_x_noised else {}
    np.savez_compressed(os.path.join(OUT, fname), seed=seed, seq_len=seq_len, mode=mode,
                        diffusion_steps=diffusion_steps, step=step, strength=strength, top_p=top_p,
                        input_ids=cond["input_ids"], input_mask=cond["input_mask"],
                        step_ids=np.stack(rec_ids), step_margin=np.stack(rec_margin),
                        **({"final_sample": sample.numpy()} if store_final else {}), tokens=tokens.numpy(),
                        logit_margin=(top2[..., 0] - top2[..., 1]).numpy().astype(np.float32), **extra)
    print(fname, tokens.shape, len(rec_ids), "rounding calls, %.0f s" % (time.time() - t0), tokens[0, :20].tolist())


def corruption_cases(n=40, L=320, seed=4):
    """rows shaped like the modification pipeline's input (`[meta .. EOS notes .. EOS pad]`, data/preprocess.py:50-56)."""
    rows, r = [], 0
    while len(rows) < n:
        row = O.make_synthetic_batch("modification", 1, L, seed=seed * 100 + r)["input_ids"][0].astype(np.int64)
        r += 1
        bars, eos = np.flatnonzero(row == 2), np.flatnonzero(row == 1)
        if len(bars) >= 3 and eos[-1] > bars[-1]:   # random_rotating asserts >= 2 bars and cuts the last bar at the final EOS
            rows.append(row)
    return np.stack(rows)


def golden_corruption():
    """the reference's own Corruptions (data/corruption.py) on 40 rows, one seeded stream:
    each single corruption, the default config `mt,mn,rn,rr` / corr_max 4 / corr_p 0.5, and a config with kwargs."""
    import torch
    from MuseDiffusion.data import corruption as RC
    rows = corruption_cases()
    out = {"rows": rows}
    configs = {"mt": ("mt", 1, 1.0, None), "mn": ("mn", 1, 1.0, None), "rn": ("rn", 1, 1.0, None), "rr": ("rr", 1, 1.0, None),
               "default": ("mt,mn,rn,rr", 4, 0.5, None), "kw": ("rr,mt,rn", 2, 0.7, "dict(p=0.15, count=2)")}
    for tag, (avail, cmax, p, kw) in configs.items():
        corr = RC.Corruptions.from_config(avail, cmax, p, kw)
        RC.generator.seed(1234)
        res = [corr(torch.from_numpy(r)).numpy() for r in rows]
        # a rotation after a masked final EOS changes the row length (corruption.py:185-193): store padded with -1 + lengths
        width = max(len(x) for x in res)
        out[tag] = np.stack([np.concatenate([x, np.full(width - len(x), -1, np.int64)]) for x in res])
        out[tag + "_len"] = np.array([len(x) for x in res], np.int64)
        out[tag + "_next"] = np.float64(RC.generator.random())          # the stream position afterwards
    np.savez_compressed(os.path.join(OUT, "corruption.npz"), **out)
    print("corruption.npz", {k: (int(out[k + "_len"].min()), int(out[k + "_len"].max())) for k in configs})


def golden_training_args():
    """`training_args.json` exactly as the reference's trainer writes it
    (`TrainSettings(...).json()`, pydantic v1; config/train.py:101-124) — defaults, and one with non-default model fields."""
    from MuseDiffusion.config import TrainSettings
    with open(os.path.join(OUT, "training_args_default.json"), "w") as f:
        f.write(TrainSettings().json())
    with open(os.path.join(OUT, "training_args_small.json"), "w") as f:
        f.write(TrainSettings(seq_len=256, diffusion_steps=400, noise_schedule="cosine", predict_xstart=False,
                              rescale_timesteps=False, timestep_respacing="ddim50", use_corruption=False).json())
    print("training_args_*.json")


def golden_meta_prefix():
    """README example meta -> 27-token prefix through the real MetaToSequence."""
    from MuseDiffusion.utils.decode_util import meta_to_batch
    meta = {"bpm": 70, "audio_key": "aminor", "time_signature": "4/4", "pitch_range": "mid_high",
            "num_measures": 8, "inst": "acoustic_piano", "genre": "newage", "min_velocity": 60,
            "max_velocity": 80, "track_role": "main_melody", "rhythm": "standard",
            "chord_progression": "-".join(["Am"] * 8 + ["G"] * 8 + ["F"] * 8 + ["E"] * 8) + "-" +
                                 "-".join(["Am"] * 8 + ["G"] * 8 + ["F"] * 8 + ["E"] * 8)}
    b = meta_to_batch(meta, batch_size=2, seq_len=64)
    np.savez_compressed(os.path.join(OUT, "meta_batch.npz"), input_ids=b["input_ids"].numpy(),
                        input_mask=b["input_mask"].numpy())
    print("meta_batch.npz", b["input_ids"][0, :30].tolist())


def meta_encode_cases(n=60, seed=9):
    """Seeded meta dicts over every category table (+ chord progressions with in-bar changes and 'NN')."""
    import random as _r
    from commu.preprocessor.utils import constants as C
    from commu.preprocessor.encoder.event_tokens import base_event
    rng = _r.Random(seed)
    chords = [k[6].upper() + k[7:] for k in base_event if k.startswith("Chord_")]
    cases = []
    for i in range(n):
        bars = rng.choice([1, 2, 4, 8, 16])
        prog, cur = [], rng.choice(chords)
        for _ in range(bars * 8):
            if rng.random() < 0.25:
                cur = rng.choice(chords)
            prog.append(cur)
        cases.append({"bpm": rng.choice([0, 3, 5, 37, 70, 120, 199, 200, 201, 400]), "audio_key": rng.choice(list(C.KEY_MAP)),
                      "time_signature": rng.choice(list(C.TIME_SIG_MAP)), "pitch_range": rng.choice(list(C.PITCH_RANGE_MAP)),
                      "num_measures": rng.choice([4, 4.5, 5, 8, 8.9, 9, 16, 17]), "inst": rng.choice(list(C.INST_MAP)),
                      "genre": rng.choice(list(C.GENRE_MAP)), "min_velocity": rng.randrange(0, 128),
                      "max_velocity": rng.randrange(0, 128), "track_role": rng.choice(list(C.TRACK_ROLE_MAP)),
                      "rhythm": rng.choice(list(C.RHYTHM_MAP)), "chord_progression": "-".join(prog)})
    return cases


def golden_meta_encode():
    """tests/golden/meta_encode.json: MetaToSequence.execute (utils/decode_util.py:44-47) on seeded meta dicts; the
    per-field "unknown" tokens and the error cases through commu's encode_meta on a plain attribute object (pydantic
    would reject "unknown" for the int fields before the encoder sees it)."""
    import json
    from types import SimpleNamespace
    from MuseDiffusion.utils.decode_util import MetaToSequence
    from commu.preprocessor.encoder.meta import encode_meta, META_ENCODING_ORDER
    from commu.preprocessor.utils.exceptions import UnprocessableMidiError
    m2s = MetaToSequence()
    cases = meta_encode_cases()
    out = {"cases": [{"meta": c, "tokens": [int(t) for t in m2s.execute(c)]} for c in cases], "unknown": [], "errors": []}
    base = {k: v for k, v in cases[0].items() if k != "chord_progression"}
    for name in META_ENCODING_ORDER:
        d = dict(base)
        d[name] = "unknown"
        try:
            out["unknown"].append({"field": name, "tokens": [int(t) for t in encode_meta(SimpleNamespace(**d))]})
        except UnprocessableMidiError:
            out["errors"].append({"field": name, "value": "unknown"})
    for name, bad in (("audio_key", "hmajor"), ("num_measures", 7), ("num_measures", 32), ("inst", "kazoo"),
                      ("time_signature", "5/4"), ("rhythm", "swing")):
        d = dict(base)
        d[name] = bad
        try:
            encode_meta(SimpleNamespace(**d))
            raise SystemExit("expected an error for %s=%s" % (name, bad))
        except UnprocessableMidiError:
            out["errors"].append({"field": name, "value": bad})
    json.dump(out, open(os.path.join(OUT, "meta_encode.json"), "w"), indent=0)
    print("meta_encode.json", len(out["cases"]), "cases,", len(out["unknown"]), "unknown,", len(out["errors"]), "errors")


def golden_decode_prepare():
    """the reference's own SequenceToMidi (decode_util.py:57-199) on the rows above."""
    from MuseDiffusion.utils.decode_util import SequenceToMidi, SequenceToMidiError
    codes = {"NO EOS TOKEN": 1, "RESTORE_CHORD FROM META FAILED": 2, "VALIDATION OF SEQUENCE FAILED": 3,
             "STRICT VALIDATION OF SEQUENCE FAILED": 4}
    from decode_oracle import decode_cases
    tokens, masks = decode_cases()
    B, L = tokens.shape
    out = {}
    for strict in (0, 1):
        dec = SequenceToMidi(strict_validation=bool(strict))
        status = np.zeros((B,), np.int32); note_len = np.zeros((B,), np.int32)
        notes = np.zeros((B, 2 * L), np.int32); meta = np.zeros((B, 11), np.int32)
        for b in range(B):
            ns = mt = None
            try:
                ns, mt = dec.split_meta_midi(tokens[b], masks[b])
                dec.validate_generated_sequence(ns)
            except SequenceToMidiError as e:
                status[b] = codes[str(e)]
            except IndexError:
                status[b] = 5
            if ns is not None:
                note_len[b] = len(ns); notes[b, :len(ns)] = ns; meta[b, :len(mt)] = mt
        out["status_%d" % strict], out["note_len_%d" % strict], out["notes_%d" % strict], out["meta_%d" % strict] = status, note_len, notes, meta
        print("decode_prepare strict=%d status histogram" % strict, np.bincount(status, minlength=6).tolist())
    np.savez_compressed(os.path.join(OUT, "decode_prepare.npz"), tokens=tokens, masks=masks, **out)


def _recording_miditoolkit():
    """miditoolkit is absent here: give the reference's write_midi (commu/preprocessor/encoder/encoder_utils.py:386-497)
    plain containers that keep what they are handed, so its event walk runs unmodified and its output can be read back."""
    import sys as _sys
    from types import SimpleNamespace as NS
    mt = _sys.modules["miditoolkit"]

    def rec(kind, *names):
        def make(*a, **k):
            d = dict(zip(names, a))
            d.update(k)
            return NS(kind=kind, **d)
        return make

    class MidiFile:
        def __init__(self, *a, **k):
            self.time_signature_changes, self.key_signature_changes, self.tempo_changes = [], [], []
            self.instruments, self.markers, self.ticks_per_beat = [], [], None
    mt.Note = rec("note", "velocity", "pitch", "start", "end")
    mt.KeySignature = rec("key", "key_name", "time")
    mt.MidiFile = MidiFile
    mt.midi.parser.MidiFile = MidiFile
    c = mt.midi.containers
    c.TimeSignature = rec("ts", "numerator", "denominator", "time")
    c.TempoChange = rec("tempo", "tempo", "time")
    c.Marker = rec("marker", "text", "time")

    def instrument(program, is_drum=False, name=""):
        return NS(kind="inst", program=program, is_drum=is_drum, name=name, notes=[])
    c.Instrument = instrument


def golden_midi_decode():
    """tests/golden/midi_decode.json: the reference's decode_event_sequence (utils/decode_util.py:201-205 ->
    EventSequenceEncoder.decode -> write_midi) on every row of decode_cases() that passes validate_once, with the OOV lines
    it prints; rows on which it raises keep the exception's class name."""
    import contextlib
    import io
    import json
    from MuseDiffusion.utils.decode_util import SequenceToMidi, SequenceToMidiError
    from decode_oracle import decode_cases
    _recording_miditoolkit()
    tokens, masks = decode_cases()
    dec = SequenceToMidi(strict_validation=False)
    rows = []
    for b in range(len(tokens)):
        try:
            ns, mt = dec.split_meta_midi(tokens[b], masks[b])
            dec.validate_generated_sequence(ns)
        except (SequenceToMidiError, IndexError):
            continue
        log = io.StringIO()
        row = {"row": b}
        try:
            with contextlib.redirect_stdout(log):
                midi = dec.decode_event_sequence(ns, mt)
        except Exception as exc:                                   # e.g. KeyError for an "unknown" key / time-signature token
            row["error"] = exc.__class__.__name__
        else:
            inst, = midi.instruments
            row.update(ticks_per_beat=int(midi.ticks_per_beat), program=int(inst.program), is_drum=bool(inst.is_drum),
                       tempo=[[int(t.tempo), int(t.time)] for t in midi.tempo_changes],
                       time_signature=[[int(t.numerator), int(t.denominator), int(t.time)] for t in midi.time_signature_changes],
                       key=[[k.key_name, int(k.time)] for k in midi.key_signature_changes],
                       notes=[[int(n.velocity), int(n.pitch), int(n.start), int(n.end)] for n in inst.notes],
                       markers=[[m.text, int(m.time)] for m in midi.markers], oov=log.getvalue().splitlines())
        rows.append(row)
    # direct cases: valid meta over every time signature / key, long note sequences, chords, OOV words, broken groups
    rng = np.random.default_rng(77)
    direct = []
    for k in range(48):
        mt = np.array([int(rng.integers(561, 601)), int(rng.integers(602, 626)), 627 + k % 4, int(rng.integers(631, 638)),
                       int(rng.integers(638, 641)), int(rng.integers(642, 650)), int(rng.integers(651, 653)),
                       int(rng.integers(654, 719)), int(rng.integers(654, 719)), int(rng.integers(720, 726)),
                       int(rng.integers(727, 729))])
        ns = []
        for _ in range(int(rng.integers(1, 9))):
            ns.append(2)
            for _ in range(int(rng.integers(0, 12))):
                u = rng.random()
                if u < 0.70:
                    ns += [int(rng.integers(432, 560)), int(rng.integers(131, 195)), int(rng.integers(3, 131)), int(rng.integers(304, 432))]
                elif u < 0.85:
                    ns += [int(rng.integers(432, 560)), int(rng.integers(195, 304))]
                elif u < 0.92:
                    ns.append(int(rng.choice([0, 560, 600, 650, 728, 1000])))          # words outside the event vocabulary
                else:
                    ns += [int(rng.integers(432, 560)), int(rng.integers(131, 195))]   # a group cut short
        if k % 3 == 0:
            ns = ns[1:]                                                                 # sequence that does not open with Bar
        ns.append(1)
        ns = np.array(ns)
        log = io.StringIO()
        with contextlib.redirect_stdout(log):
            midi = dec.decode_event_sequence(ns, mt)
        inst, = midi.instruments
        direct.append(dict(note_seq=ns.tolist(), meta=mt.tolist(), ticks_per_beat=int(midi.ticks_per_beat),
                           tempo=[[int(t.tempo), int(t.time)] for t in midi.tempo_changes],
                           time_signature=[[int(t.numerator), int(t.denominator), int(t.time)] for t in midi.time_signature_changes],
                           key=[[kk.key_name, int(kk.time)] for kk in midi.key_signature_changes],
                           notes=[[int(n.velocity), int(n.pitch), int(n.start), int(n.end)] for n in inst.notes],
                           markers=[[m.text, int(m.time)] for m in midi.markers], oov=log.getvalue().splitlines()))
    json.dump({"rows": rows, "direct": direct}, open(os.path.join(OUT, "midi_decode.json"), "w"))
    print("direct:", len(direct), "cases,", sum(len(r["notes"]) for r in direct), "notes,", sum(len(r["markers"]) for r in direct),
          "markers,", sum(len(r["oov"]) for r in direct), "OOV lines")
    print("midi_decode.json", len(rows), "rows,", sum("error" in r for r in rows), "errors,",
          sum(len(r.get("notes", ())) for r in rows), "notes,", sum(bool(r.get("oov")) for r in rows), "rows with OOV")


def golden_merge_and_mask(seq_len=96):
    """the reference's own helper_tokenize (data/preprocess.py:26-70), helper_filter
    (:73-81) and collate_batches (data/wrapper.py:90-126) on the rows of preprocess_oracle.merge_cases()."""
    import torch
    from MuseDiffusion.data.preprocess import helper_tokenize, helper_filter
    from MuseDiffusion.data.wrapper import collate_batches
    from preprocess_oracle import merge_cases
    src, src_len, trg, trg_len = merge_cases()
    B = len(src_len)
    raw = {"src": [src[b, :src_len[b]].tolist() for b in range(B)], "trg": [trg[b, :trg_len[b]].tolist() for b in range(B)]}
    merged = helper_tokenize(raw, num_proc=1)
    length = np.array(merged["length"], np.int64)
    kept = helper_filter(merged, seq_len=seq_len, num_proc=1)
    assert len(kept) == int((length <= seq_len).sum())
    rows = [{"input_ids": torch.tensor(kept["input_ids"][i]), "input_mask": torch.tensor(kept["input_mask"][i]),
             "length": kept["length"][i]} for i in range(len(kept))]
    col = collate_batches(rows, seq_len=seq_len)
    np.savez_compressed(os.path.join(OUT, "merge_and_mask.npz"), src=src, src_len=src_len, trg=trg, trg_len=trg_len,
                        seq_len=seq_len, length=length, kept_input_ids=col["input_ids"].numpy(),
                        kept_input_mask=col["input_mask"].numpy(), kept_length=col["length"].numpy())
    print("merge_and_mask.npz", B, "rows,", len(kept), "kept at seq_len", seq_len)


def golden_metrics():
    """the reference's own metric.py (get_vectors, ONNC, Controllability_*) on the
    sequences of metric_oracle.metric_cases()."""
    import torch
    from MuseDiffusion.metric import get_vectors, ONNC, Controllability_Pitch, Controllability_Velocity
    from metric_oracle import metric_cases
    metas, midis, lens = metric_cases()
    rows = [midis[b, :lens[b]] for b in range(len(lens))]
    vecs = [get_vectors(r) for r in rows]
    rhythm = torch.stack([v[0] for v in vecs]).numpy()
    melody = torch.stack([v[1] for v in vecs]).numpy()
    harmony = torch.stack([v[2] for v in vecs]).numpy()
    score, msim, most = ONNC(rows, return_MSIM=True, return_mostsim=True)
    cp = Controllability_Pitch(metas, rows)
    cv = Controllability_Velocity(metas, rows)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), metas=metas, midis=midis, lens=lens, rhythm=rhythm, melody=melody,
                        harmony=harmony, onnc=float(score), msim=msim.numpy(), most_sim=most.numpy(),
                        cp=np.array(cp, np.int64), cv=np.array(cv, np.int64))
    print("metrics.npz", len(rows), "rows, ONNC %.4f CP %s CV %s" % (float(score), cp, cv))


def main():
    os.makedirs(OUT, exist_ok=True)
    install_reference_shim()
    if len(sys.argv) > 1 and sys.argv[1] == "metrics":
        golden_metrics()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "meta":
        golden_meta_encode()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "decode":
        golden_decode_prepare()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "midi":
        golden_midi_decode()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "merge":
        golden_merge_and_mask()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "corruption":
        golden_corruption()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "args":
        golden_training_args()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "loops":
        import torch
        torch.manual_seed(0)
        torch.set_num_threads(os.cpu_count())
        golden_loop("loop_gen_ddpm.npz", "generation", 64, 2, 11, 40, 40)
        golden_loop("loop_mod_ddpm.npz", "modification", 64, 3, 12, 40, 40, strength=0.75)
        golden_loop("loop_mod_ddim.npz", "modification", 64, 2, 13, 2000, 20, strength=1.0)
        golden_loop("loop_gen_ddim.npz", "generation", 96, 2, 14, 2000, 10)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "base":
        # ~15 min on 8 cores: BASELINE.json configs[0] (modification, --step 100 => DDIM gap 20, strength 1.0) and a
        # generation-mode DDPM chain (truncated noise, clamp every step) at the base sequence length
        import torch
        torch.manual_seed(0)
        torch.set_num_threads(os.cpu_count())
        golden_loop_base("loop_base_gen_ddpm24.npz", "generation", 2, 21, 24, 24)
        golden_loop_base("loop_base_mod_ddim100.npz", "modification", 2, 22, 2000, 100, strength=1.0)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "base_b4":
        # BASELINE.json configs[0] at its own batch size: 4 sequences x 2096, modification, --step 100 (DDIM gap 20), strength 1.0;
        # ~15 min on 8 cores.  Ids + margins only (the B = 2 fixture above also pins the final sample).
        import torch
        torch.manual_seed(0)
        torch.set_num_threads(os.cpu_count())
        golden_loop_base("loop_base_mod_ddim100_b4.npz", "modification", 4, 24, 2000, 100, strength=1.0, store_x_noised=False,
                         store_final=False)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "base_ddpm":
        # the bench's own operating regime (BASELINE.json configs[1]): modification, 2000-step table, DDPM with truncated noise,
        # rounding every step, the first 40 indices from the top of the chain (strength 0.02 -> t_enc = 40); ~4 min on 8 cores
        import torch
        torch.manual_seed(0)
        torch.set_num_threads(os.cpu_count())
        golden_loop_base("loop_base_mod_ddpm40.npz", "modification", 2, 23, 2000, 2000, strength=0.02, store_x_noised=False,
                         store_final=False)
        return
    import torch
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    golden_schedules()
    golden_rounding()
    golden_meta_prefix()
    golden_meta_encode()
    golden_forward(64, 2, 1, "forward_tiny.npz", [999.5, 3.0])
    golden_forward(200, 1, 2, "forward_ragged.npz", [500.0])          # L not a multiple of 64/128
    golden_forward(2096, 1, 4, "forward_base.npz", [250.0])           # the base-config shape
    golden_steps()
    golden_loop("loop_gen_ddpm.npz", "generation", 64, 2, 11, 40, 40)
    golden_loop("loop_mod_ddpm.npz", "modification", 64, 3, 12, 40, 40, strength=0.75)
    golden_loop("loop_mod_ddim.npz", "modification", 64, 2, 13, 2000, 20, strength=1.0)
    golden_loop("loop_gen_ddim.npz", "generation", 96, 2, 14, 2000, 10)
    golden_decode_prepare()
    golden_midi_decode()
    golden_merge_and_mask()
    golden_metrics()
    golden_training_args()
    golden_corruption()


if __name__ == "__main__":
    main()
