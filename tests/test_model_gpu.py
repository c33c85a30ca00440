"""End-to-end parity of the CUDA model against the golden outputs of the unmodified reference
(tests/golden/case_*.npz) and, at larger shapes, against the oracle run on the box's CPU."""
import numpy as np
import pytest
import torch

from helpers import CASE_NAMES, case_batch, case_config, case_state_dict, load_case, lookup
from test_ops_gpu import close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_TOL = 1e-4  # north_star: fp32 rel ≤ 1e-4


def build_model(meta):
    from fastspeech2_lightning_b200 import synthetic
    from fastspeech2_lightning_b200.fs2.model import FastSpeech2

    model = FastSpeech2(case_config(meta), stats=synthetic.DEFAULT_STATS,
                        lang2id=lookup(meta.get("n_languages", 0), "l"), speaker2id=lookup(meta.get("n_speakers", 0), "s"))
    missing = model.load_state_dict(case_state_dict(meta), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.postnet.dropout_in_training = False  # parity runs: PostNet's hard-coded dropout off
    model = model.to(DEV)
    model.train(meta["mode"] != "eval")
    model.current_epoch = meta.get("epoch", 0)
    return model


@pytest.mark.parametrize("name", [n for n in CASE_NAMES])
def test_forward_matches_reference_golden(name):
    meta, gold = load_case(name)
    model = build_model(meta)
    batch = case_batch(meta, DEV)
    with torch.no_grad():
        out = model(batch, inference=meta["inference"])
    for k, want in gold.items():
        if not k.startswith("out."):
            continue
        got = out[k[4:]]
        assert got is not None, k
        if want.dtype.kind in "biu":  # masks, lengths, durations: bit exact
            assert np.array_equal(got.cpu().numpy().astype(want.dtype), want), k
        elif k == "out.attn_hard":
            assert np.array_equal(got.cpu().numpy(), want), k
        else:
            close(got, want, FP32_TOL, f"{name}:{k}")
    if not meta["inference"]:
        losses = model.loss(out, batch, model.current_epoch)
        for k, want in gold.items():
            if k.startswith("loss."):
                close(losses[k[5:]].detach().cpu().reshape(()), np.asarray(want, dtype=np.float64), FP32_TOL, f"{name}:{k}")


def test_output_dict_contract():
    meta, _ = load_case("train_eval")
    model = build_model(meta)
    out = model(case_batch(meta, DEV)) if False else None
    with torch.no_grad():
        out = model(case_batch(meta, DEV))
    assert list(out.keys()) == [
        "output", "postnet_output", "src_mask", "src_lens", "tgt_mask", "tgt_lens", "attn_logprob", "attn_soft",
        "attn_hard", "duration_prediction", "duration_target", "energy_prediction", "energy_target",
        "pitch_prediction", "pitch_target", "text_input"]
    assert out["src_mask"].dtype == torch.bool and out["tgt_mask"].dtype == torch.bool
    assert out["duration_target"].dtype == torch.int32
    assert out["attn_hard"].shape == out["attn_soft"].shape == out["attn_logprob"].shape


def test_c1_shape_against_oracle_on_cpu():
    """BASELINE configs[0]: B=16, T≈80, F≈500 teacher-forced synthesis, vs the oracle on the host CPU."""
    from fastspeech2_lightning_b200 import synthetic
    from fastspeech2_lightning_b200.fs2.config import FastSpeech2Config
    from fastspeech2_lightning_b200.fs2.model import FastSpeech2
    from oracle import fs2_oracle

    cfg = FastSpeech2Config(model=dict(learn_alignment=False))
    torch.manual_seed(0)
    model = FastSpeech2(cfg, stats=synthetic.DEFAULT_STATS)
    synthetic.fill_weights_(model, seed=21)
    model.eval()
    batch = synthetic.make_batch(16, (60, 80), seed=5, learn_alignment=False, inference=True, teacher_forced=True)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        want = fs2_oracle.forward(sd, fs2_oracle.Cfg(cfg), batch, inference=True)
        got = model.to(DEV)(synthetic.batch_to(batch, DEV), inference=True)
    assert torch.equal(got["tgt_mask"].cpu(), want["tgt_mask"])
    close(got["output"], want["output"], FP32_TOL, "C1 output")
    close(got["postnet_output"], want["postnet_output"], FP32_TOL, "C1 postnet_output")
    close(got["duration_prediction"], want["duration_prediction"], FP32_TOL, "C1 log-dur")
