"""Shared test helpers: load golden cases, rebuild their inputs from the seeds."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from fastspeech2_lightning_b200 import synthetic
from fastspeech2_lightning_b200.fs2.config import FastSpeech2Config

GOLDEN = Path(__file__).resolve().parent / "golden"
CASE_NAMES = sorted(p.stem[5:] for p in GOLDEN.glob("case_*.json"))


def load_case(name):
    meta = json.loads((GOLDEN / f"case_{name}.json").read_text())
    arrays = dict(np.load(GOLDEN / f"case_{name}.npz", allow_pickle=False))
    return meta, arrays


def case_config(meta) -> FastSpeech2Config:
    m = json.loads(json.dumps(meta["model"]))
    if meta["mode"] == "train_nodrop":
        m.setdefault("encoder", {})["dropout"] = 0.0
        m.setdefault("decoder", {})["dropout"] = 0.0
        vp = m.setdefault("variance_predictors", {})
        for k in ("energy", "pitch", "duration"):
            vp.setdefault(k, {})["dropout"] = 0.0
    return FastSpeech2Config(model=m)


def case_batch(meta, device="cpu"):
    b = synthetic.make_batch(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in meta["batch"].items()})
    if meta.get("frame_energy"):
        g = np.random.default_rng(meta["batch"]["seed"] + 1000)
        F = int(b["max_mel_len"])
        valid = np.arange(F)[None, :] < b["mel_lens"].numpy()[:, None]
        b["energy"] = torch.from_numpy((g.standard_normal((len(valid), F)) * valid).astype(np.float32))
    return synthetic.batch_to(b, device)


def case_state_dict(meta):
    """The synthetic state dict of a golden case, rebuilt from its seed (no reference needed)."""
    sd = {}
    st = synthetic.DEFAULT_STATS
    for k, shape in meta["state_shapes"].items():
        v = synthetic.synth_tensor(k, shape, meta["weights_seed"])
        if v is None:
            leaf = k.rsplit(".", 1)[-1]
            if leaf == "inv_freq":
                v = 1 / (10000 ** (torch.arange(0.0, 256, 2.0) / 256))
            else:
                s = st["pitch" if "pitch" in leaf else "energy"]
                v = torch.linspace(s["norm_min"], s["norm_max"], shape[0])
        if k == "text_input_layer.weight" and v.shape[1] != 39:
            v[0].zero_()
        sd[k] = v
    if "dur_bias" in meta:
        sd["variance_adaptor.duration_predictor.linear.bias"].fill_(meta["dur_bias"])
    return sd


def lookup(n, prefix):
    return {f"{prefix}{i}": i for i in range(n)}
