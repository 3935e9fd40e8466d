"""SURVEY.md section 8(f) row 3 — on-disk formats: a model directory as the reference's trainer writes it
(`model_*.pt` + `TrainSettings(...).json()`; the two json fixtures were written by the UNMODIFIED reference through
oracle/make_golden.py::golden_training_args) loads without blobfile / pydantic, and the packed weight file round-trips
bit for bit and drives the kernels to the same output as the state dict it was written from."""
import json
import os

import numpy as np
import pytest
import torch

import musediff_oracle as O
from musediffusion_b200 import checkpoint as C
from musediffusion_b200.initialization import create_model_and_diffusion
from musediffusion_b200.sample import load_model, main as cli_main


def test_reference_written_training_args_load(golden_dir):
    a = C.load_training_args(os.path.join(golden_dir, "training_args_default.json"))
    assert (a.seq_len, a.vocab_size, a.hidden_dim, a.hidden_t_dim, a.diffusion_steps) == (2096, 729, 128, 128, 2000)
    assert a.noise_schedule == "sqrt" and a.predict_xstart is True and a.rescale_timesteps is True and a.timestep_respacing == ""
    assert a.use_corruption is True and a.corr_available == "mt,mn,rn,rr"          # carried along untouched
    b = C.load_training_args(os.path.join(golden_dir, "training_args_small.json"))
    assert (b.seq_len, b.diffusion_steps, b.noise_schedule, b.predict_xstart, b.rescale_timesteps, b.timestep_respacing) == \
        (256, 400, "cosine", False, False, "ddim50")
    model, diffusion = create_model_and_diffusion(b)
    assert model.position_embeddings.weight.shape[0] == 256 and diffusion.num_timesteps == 50 and not diffusion.predict_xstart
    assert len(model.state_dict()) == 211


def test_model_directory_and_pack_roundtrip(tmp_path, golden_dir):
    d = tmp_path / "diffusion_models"
    d.mkdir()
    raw = json.load(open(os.path.join(golden_dir, "training_args_default.json")))
    raw["seq_len"] = 64
    (d / "training_args.json").write_text(json.dumps(raw))
    p = O.make_random_params(seed=9, seq_len=64)
    sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in p.items()}
    torch.save(sd, str(d / "model_000100.pt"))
    # the directory / the checkpoint / the json all resolve to the same settings (config/sample.py:123-134)
    for where in (str(d), str(d / "model_000100.pt"), str(d / "training_args.json")):
        assert C.load_training_args(where).seq_len == 64
    back = C.load_state_dict(str(d / "model_000100.pt"))
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    # pack through the CLI (host only), read back: bit-identical to packing in memory, 256-byte aligned views of one buffer
    out = cli_main(["pack", "--model_path", str(d / "model_000100.pt")])
    assert out.endswith("model_000100.mdpack") and os.path.getsize(out) < 0.51 * os.path.getsize(str(d / "model_000100.pt"))
    cfg, tensors = C.load_pack(out, "cpu")
    ref = C.pack_tensors(sd, 12)
    assert list(tensors) == list(ref)
    for k in ref:
        assert tensors[k].dtype == ref[k].dtype and tensors[k].shape == ref[k].shape
        assert torch.equal(tensors[k].view(torch.uint8), ref[k].view(torch.uint8)), k
        assert tensors[k].data_ptr() % 256 == 0
    assert cfg["seq_len"] == 64 and cfg["num_hidden_layers"] == 12 and cfg["diffusion_steps"] == 2000
    assert tensors["l3.wqkv"].shape == (2304, 768) and tensors["l3.wqkv"].dtype == torch.bfloat16
    # 1/sqrt(64) folded into the query rows exactly
    q = sd["input_transformers.layer.3.attention.self.query.weight"]
    assert torch.equal(tensors["l3.wqkv"][:768].float(), (q * 0.125).to(torch.bfloat16).float())
    with pytest.raises(ValueError):
        C.read_pack_header(str(d / "model_000100.pt"))


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")
def test_packed_weights_drive_the_same_forward(tmp_path, golden_dir):
    dev = torch.device("cuda:0")
    d = tmp_path / "m"
    d.mkdir()
    raw = json.load(open(os.path.join(golden_dir, "training_args_default.json")))
    raw["seq_len"] = 200
    (d / "training_args.json").write_text(json.dumps(raw))
    p = O.make_random_params(seed=2, seq_len=200)
    torch.save({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in p.items()}, str(d / "model_000000.pt"))
    pack = cli_main(["pack", "--model_path", str(d / "model_000000.pt")])
    m1, diff1, _ = load_model(str(d / "model_000000.pt"), dev)
    m2, diff2, _ = load_model(pack, dev)
    g = np.load(os.path.join(golden_dir, "forward_ragged.npz"))           # reference forward at this seed / length
    x = torch.from_numpy(g["x"]).to(dev)
    t = torch.from_numpy(g["t"]).to(dev)
    y1, y2 = m1(x, t), m2(x, t)
    assert torch.equal(y1, y2)                                            # same packed bits -> same kernels -> same output
    ref = g["model_output"]
    assert float(np.abs(y2.cpu().numpy() - ref).max() / np.abs(ref).max()) < 2e-2
    assert np.array_equal(m2.decode_tokens(x).cpu().numpy(), g["tokens"])
    assert torch.equal(m2.word_embedding.weight, m1.word_embedding.weight) and torch.equal(m2.lm_head.bias, m1.lm_head.bias)
    assert diff1.num_timesteps == diff2.num_timesteps == 2000
    # load_state_dict afterwards takes the parameters again
    p2 = O.make_random_params(seed=5, seq_len=200)
    m2.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in p2.items()})
    assert not torch.equal(m2(x, t), y1)


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")
def test_cli_samples_from_checkpoint_and_from_pack(tmp_path, golden_dir):
    """`python -m musediffusion_b200 modification|generation` (flag names of config/sample.py:93-209) on a reference-style model
    directory and on the pack written from it: same seed -> same decoded ids; encoder passes of any size give the same bits."""
    d = tmp_path / "m"
    d.mkdir()
    raw = json.load(open(os.path.join(golden_dir, "training_args_default.json")))
    raw["seq_len"] = 64
    (d / "training_args.json").write_text(json.dumps(raw))
    p = O.make_random_params(seed=12, seq_len=64)
    torch.save({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in p.items()}, str(d / "model_000000.pt"))
    pack = cli_main(["pack", "--model_path", str(d / "model_000000.pt")])
    outs = []
    for mp, sub in ((str(d / "model_000000.pt"), "a"), (pack, "b")):
        cli_main(["modification", "--model_path", mp, "--step", "10", "--batch_size", "3", "--num_batches", "2", "--strength", "1.0",
                  "--out_dir", str(tmp_path / sub)])
        base = tmp_path / sub / "m" / (os.path.basename(mp) + ".modification.samples")
        tok = np.load(str(base / "tokens.npy"))
        st = np.load(str(base / "decode_status.npy"))
        assert tok.shape == (6, 64) and tok.dtype == np.int64 and st.shape == (6,)
        outs.append(tok)
    assert np.array_equal(outs[0], outs[1])
    cli_main(["generation", "--model_path", pack, "--step", "20", "--batch_size", "2", "--num_samples", "4", "--out_dir", str(tmp_path / "g")])
    gtok = np.load(str(tmp_path / "g" / "m" / (os.path.basename(pack) + ".generation.samples") / "tokens.npy"))
    assert gtok.shape == (4, 64)
    # micro-batched encoder passes == one pass, bit for bit
    dev = torch.device("cuda:0")
    m, _, _ = load_model(pack, dev)
    x = torch.randn(5, 64, 128, device=dev)
    t = torch.tensor([10.0, 500.0, 3.0, 999.0, 77.0], device=dev)
    one = m.denoise(x, t).clone()
    m.max_tokens_per_pass = 2 * 64
    assert m.pass_size(5, 64) == 2
    assert torch.equal(m.denoise(x, t), one)
    assert torch.equal(m.denoise(x, t[:1].expand(5).contiguous(), uniform_t=True), m.denoise(x, t[:1].expand(5).contiguous()))
