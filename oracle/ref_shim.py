"""ORACLE-side test infrastructure (not product code): import the UNMODIFIED reference package `fs2`.

The sources are taken from `/root/reference` when it exists (the build container) and otherwise from `oracle/_ref/`,
a build-time copy made by `oracle/build_ref.py` (git-ignored; it travels to the GPU box with the snapshot the way
the built `.so` files do).  Used by `tests/golden/make_golden.py`, the oracle pin tests and `bench.py`'s reference arm /
`cpu_baseline` (`oracle/ref_runner.py`).  The reference depends on packages that are not installed and cannot
be (no network): `everyvoice`, `pytorch_lightning`, `matplotlib`.  None of them
does arithmetic on the hot path, so they are replaced by inert stubs registered
in `sys.modules` *before* `import fs2.model`; `fs2.config` (whose real schema
needs everyvoice base classes) is replaced by this repo's stand-alone schema,
which carries the same attribute paths and defaults.  Everything numerical —
`fs2/model.py`, `variance_adaptor.py`, `layers.py`, `blocks.py`, `attn/*`,
`loss.py`, `gst/*`, `torchaudio.models.Conformer`, numba MAS — is then the
reference's own code.
"""
from __future__ import annotations

import enum
import sys
import types
from pathlib import Path

REPO_ROOT = Path(__file__).resolve().parents[1]
_CANDIDATES = (Path("/root/reference"), REPO_ROOT / "oracle" / "_ref")


def reference_root():
    for c in _CANDIDATES:
        if (c / "fs2" / "model.py").exists():
            return c
    return None


REFERENCE_ROOT = reference_root()


def reference_available() -> bool:
    return reference_root() is not None


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave as a package so sub-imports resolve
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def install() -> None:
    """Register the stubs and put /root/reference on sys.path (idempotent)."""
    if "fs2" in sys.modules and getattr(sys.modules["fs2"], "_is_reference", False):
        return
    import torch

    if str(REPO_ROOT) not in sys.path:
        sys.path.insert(0, str(REPO_ROOT))
    from fastspeech2_lightning_b200.fs2 import config as our_config

    # ---- pytorch_lightning ------------------------------------------------
    class LightningModule(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.current_epoch = 0
            self.global_step = 0
            self.logger = None

        def save_hyperparameters(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    _mod("pytorch_lightning", LightningModule=LightningModule)

    # ---- matplotlib -------------------------------------------------------
    _mod("matplotlib")
    _mod("matplotlib.pyplot")

    # ---- everyvoice -------------------------------------------------------
    class BadDataError(Exception):
        pass

    class TextProcessor:
        """pad symbol is id 0; the rest of the table follows the config's symbols."""

        _pad_symbol = "\x80"

        def __init__(self, text_config):
            syms = []
            for v in text_config.symbols.values():
                syms.extend(v)
            self.symbols = [self._pad_symbol] + sorted(set(syms))

        def encode_text(self, text):
            return [self.symbols.index(text)]

    _mod("everyvoice")
    _mod("everyvoice.exceptions", BadDataError=BadDataError)
    _mod("everyvoice.config")
    _mod(
        "everyvoice.config.type_definitions",
        TargetTrainingTextRepresentationLevel=our_config.TargetTrainingTextRepresentationLevel,
    )
    _mod("everyvoice.model")
    _mod("everyvoice.model.feature_prediction")
    _mod("everyvoice.model.feature_prediction.config", FeaturePredictionConfig=our_config.FastSpeech2Config)
    _mod("everyvoice.model.vocoder")
    _mod("everyvoice.model.vocoder.HiFiGAN_iSTFT_lightning")
    _mod("everyvoice.model.vocoder.HiFiGAN_iSTFT_lightning.hfgl")
    _mod(
        "everyvoice.model.vocoder.HiFiGAN_iSTFT_lightning.hfgl.utils",
        load_hifigan_from_checkpoint=None,
        synthesize_data=None,
    )
    _mod("everyvoice.text")
    _mod("everyvoice.text.features", N_PHONOLOGICAL_FEATURES=our_config.N_PHONOLOGICAL_FEATURES)
    _mod("everyvoice.text.lookups", LookupTable=dict)
    _mod("everyvoice.text.text_processor", TextProcessor=TextProcessor)
    _mod("everyvoice.text.utils", get_symbols_from_checkpoint_symbol_dict=None, symbol_sorter=None)
    _mod("everyvoice.utils", pydantic_validation_error_shortener=str, slugify=str)
    _mod("everyvoice.utils.heavy", expand=None)

    # ---- the reference package itself, with fs2.config swapped -----------
    for k in [k for k in sys.modules if k == "fs2" or k.startswith("fs2.")]:
        del sys.modules[k]
    root = reference_root()
    if root is None:
        raise ImportError("the reference sources are neither at /root/reference nor under oracle/_ref (run __graft_entry__.build())")
    sys.path.insert(0, str(root))
    import fs2  # noqa: E402  (the reference's package)

    fs2._is_reference = True
    cfg = types.ModuleType("fs2.config")
    cfg.__dict__.update({k: v for k, v in our_config.__dict__.items() if not k.startswith("__")})
    sys.modules["fs2.config"] = cfg
    fs2.config = cfg


def reference_modules():
    """Returns the reference's modules after install()."""
    install()
    import fs2.attn.alignment as alignment
    import fs2.attn.attention as attention
    import fs2.attn.attention_loss as attention_loss
    import fs2.layers as layers
    import fs2.loss as loss
    import fs2.model as model
    import fs2.variance_adaptor as variance_adaptor

    return types.SimpleNamespace(
        model=model,
        variance_adaptor=variance_adaptor,
        layers=layers,
        loss=loss,
        alignment=alignment,
        attention=attention,
        attention_loss=attention_loss,
    )


class _Dummy(enum.Enum):
    pass
