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


def gst_reference_forward(enc, speech):
    """Plain-torch statement of the reference StyleEncoder.forward (gst/model.py:87-100,179-199,241-257; gst/attn.py:172-194)
    over the parameter-holding torch layers of our module (any dtype / device torch supports) — the checker of the GST
    kernels; the product module itself has no library path."""
    import math

    B = speech.size(0)
    hs = enc.ref_enc.convs(speech.unsqueeze(1)).transpose(1, 2)
    hs = hs.contiguous().view(B, hs.size(1), -1)
    _, ref_embs = enc.ref_enc.gru(hs)
    ref = ref_embs[-1]
    m = enc.stl.mha
    tokens = torch.tanh(enc.stl.gst_embs).unsqueeze(0).expand(B, -1, -1)
    q = m.linear_q(ref.unsqueeze(1)).view(B, -1, m.h, m.d_k).transpose(1, 2)
    k = m.linear_k(tokens).view(B, -1, m.h, m.d_k).transpose(1, 2)
    v = m.linear_v(tokens).view(B, -1, m.h, m.d_k).transpose(1, 2)
    attn = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(m.d_k), dim=-1)
    x = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, -1, m.h * m.d_k)
    return m.linear_out(x).squeeze(1)
