"""
ORACLE — TEST INFRASTRUCTURE ONLY (build-container only).

Import shim that lets the UNMODIFIED reference (`/root/reference`, read-only, Python) import in
this container so that `oracle/make_golden.py` can generate the fixtures under `tests/golden/`.
`/root/reference` does not exist on the GPU box; there the verbatim copy `oracle/_ref/` (oracle/build_ref.sh, git-ignored)
is used by the drop-in test and by bench.py's reference arm.

Why each patch is needed is recorded in SURVEY.md §8c / Appendix A:
  * `AutoConfig.from_pretrained('bert-base-uncased')` (MuseDiffusion/models/network.py:44) needs the
    network; `BertConfig()` defaults are exactly bert-base-uncased.
  * MuseDiffusion/config/base.py:7-8 uses the pydantic-v1 API; the image has pydantic 2.
  * blobfile / miditoolkit / logger / parmap / pretty_midi / yacs are absent -> stub modules.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# the build container has the read-only tree; the GPU box only has the verbatim copy made by oracle/build_ref.sh
REFERENCE_ROOT = os.environ.get("MUSEDIFFUSION_REFERENCE") or (
    "/root/reference" if os.path.isdir("/root/reference/MuseDiffusion") else os.path.join(_HERE, "_ref"))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "MuseDiffusion"))


def install_reference_shim(num_layers=None):
    """Make `import MuseDiffusion` work.  `num_layers`/other BertConfig overrides are NOT applied here;
    the golden script patches the config explicitly where it wants a smaller encoder."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torch  # noqa: F401  real imports FIRST
    import transformers  # noqa: F401
    from transformers import AutoConfig, BertConfig
    from transformers.models.bert.modeling_bert import BertEncoder  # noqa: F401  force lazy import

    def _cfg(name, **kw):
        cfg = BertConfig()
        if num_layers is not None:
            cfg.num_hidden_layers = num_layers
        return cfg

    AutoConfig.from_pretrained = staticmethod(_cfg)
    for n in ["miditoolkit", "miditoolkit.midi", "miditoolkit.midi.parser", "miditoolkit.midi.containers",
              "logger", "parmap", "pretty_midi", "yacs", "yacs.config", "blobfile"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    mt = sys.modules["miditoolkit"]
    mt.MidiFile = object
    mt.midi = sys.modules["miditoolkit.midi"]
    mt.midi.parser = sys.modules["miditoolkit.midi.parser"]
    mt.midi.containers = sys.modules["miditoolkit.midi.containers"]
    for c in ["Marker", "Instrument", "TempoChange", "Note", "TimeSignature"]:
        setattr(mt.midi.containers, c, object)
    sys.modules["logger"].logger = types.SimpleNamespace(info=print, warning=print, error=print)
    bf = sys.modules["blobfile"]
    bf.BlobFile = lambda p, m="rb": open(p, m)
    bf.join, bf.dirname, bf.exists = os.path.join, os.path.dirname, os.path.exists
    import pydantic.v1 as pv1
    import pydantic.v1.validators as pvv
    saved = {k: sys.modules.get(k) for k in ("pydantic", "pydantic.validators")}
    sys.modules["pydantic"], sys.modules["pydantic.validators"] = pv1, pvv
    try:
        import MuseDiffusion.config  # noqa: F401  the pydantic-v1 users
        import MuseDiffusion.utils.decode_util  # noqa: F401
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
