"""fastspeech2_lightning_b200 — B200-native (sm_100a) acoustic-model hot path of
EveryVoiceTTS/FastSpeech2_lightning behind the reference's own module API.

    from fastspeech2_lightning_b200.fs2.model import FastSpeech2

`install_as_fs2()` aliases the mirror package as top-level `fs2`, so code written against the
reference (`from fs2.model import FastSpeech2`, `from fs2.attn.alignment import mas_width1`, …)
picks up this implementation unchanged.
"""
import sys

__version__ = "0.1.0"


def install_as_fs2() -> None:
    import importlib

    pkg = importlib.import_module(__name__ + ".fs2")
    sys.modules["fs2"] = pkg
    for name in ("model", "variance_adaptor", "layers", "blocks", "loss", "noam", "config", "type_definitions_heavy",
                 "conformer", "attn", "attn.alignment", "attn.attention", "attn.attention_loss", "utils", "utils.heavy",
                 "gst", "gst.model"):
        sys.modules["fs2." + name] = importlib.import_module(f"{__name__}.fs2.{name}")
