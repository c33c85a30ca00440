"""The oracle (`oracle/`) pinned against outputs of the unmodified reference (tests/golden/)."""
import numpy as np
import pytest
import torch

from helpers import CASE_NAMES, GOLDEN, case_batch, case_config, case_state_dict, load_case
from oracle import fs2_oracle, intops

KATS = dict(np.load(GOLDEN / "kats.npz", allow_pickle=False))
N_MAS = len([k for k in KATS if k.startswith("mas.in.")])


@pytest.mark.parametrize("i", range(N_MAS))
def test_mas_c_and_python_restatements(i):
    x, want = KATS[f"mas.in.{i}"], KATS[f"mas.out.{i}"]
    assert np.array_equal(intops.mas_width1(x), want)
    if x.size <= 4000:
        assert np.array_equal(intops.mas_width1_py(x), want)


def test_b_mas():
    got = intops.b_mas(KATS["bmas.in"], KATS["bmas.in_lens"], KATS["bmas.out_lens"])
    assert np.array_equal(got, KATS["bmas.out"])
    # every valid frame is assigned to exactly one phone, nothing outside the window
    for b, (il, ol) in enumerate(zip(KATS["bmas.in_lens"], KATS["bmas.out_lens"])):
        assert np.all(got[b, 0, :ol, :il].sum(1) == 1)
        assert got[b, 0, ol:].sum() == 0 and got[b, 0, :, il:].sum() == 0


@pytest.mark.parametrize("tag", "abcd")
def test_length_regulator(tag):
    out, mask, idx = intops.length_regulator(KATS[f"lr.{tag}.x"], KATS[f"lr.{tag}.d"], int(KATS[f"lr.{tag}.maxlen"]))
    assert np.array_equal(out, KATS[f"lr.{tag}.out"])
    assert np.array_equal(mask, KATS[f"lr.{tag}.mask"])


def test_bucketize():
    assert np.array_equal(intops.bucketize(KATS["bucket.v"], KATS["bucket.bins"]), KATS["bucket.ids"])


@pytest.mark.parametrize("tag", "ab")
def test_average_variance(tag):
    var, dur, want = KATS[f"avg.{tag}.var"], KATS[f"avg.{tag}.dur"], KATS[f"avg.{tag}.out"]
    # fp32 prefix sums are order dependent: same torch ops → bit-equal; the
    # sequential-order numpy variant agrees to a few ulp of the running sum
    got_t = fs2_oracle.average_variance(torch.from_numpy(var), torch.from_numpy(dur)).numpy()
    assert np.array_equal(got_t, want)
    np.testing.assert_allclose(intops.average_variance(var, dur), want, rtol=0, atol=2e-6)
    assert np.array_equal(want == 0, intops.average_variance(var, dur) == 0)


@pytest.mark.parametrize("tag", "abc")
def test_round_durations(tag):
    got = intops.round_durations(KATS[f"round.{tag}.in"], float(KATS[f"round.{tag}.control"]))
    assert np.array_equal(got, KATS[f"round.{tag}.out"])


def test_positional_embedding():
    got = fs2_oracle.positional_embedding(8200, torch.from_numpy(KATS["posenc.inv_freq"]))[0, ::41].numpy()
    np.testing.assert_allclose(got, KATS["posenc.out"], rtol=0, atol=1e-6)


def _close(got, want, tol, what):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    fin = np.isfinite(want)
    assert np.array_equal(fin, np.isfinite(got)), what
    scale = max(1.0, float(np.abs(want[fin]).max())) if fin.any() else 1.0
    err = float(np.abs(got[fin] - want[fin]).max()) / scale if fin.any() else 0.0
    assert err <= tol, f"{what}: rel-to-max err {err:.3e} > {tol}"


@pytest.mark.parametrize("name", CASE_NAMES)
def test_full_model_oracle_matches_reference(name):
    meta, gold = load_case(name)
    cfg = fs2_oracle.Cfg(case_config(meta))
    sd = case_state_dict(meta)
    assert len(sd) == meta["n_state_dict"]
    batch = case_batch(meta)
    training = meta["mode"] != "eval"
    grads = bool(meta.get("grads"))
    if grads:
        for k, v in sd.items():
            if v.is_floating_point() and not k.endswith(("running_mean", "running_var", "_bins", "inv_freq")):
                v.requires_grad_(True)
    new_stats = {}
    ctx = torch.enable_grad() if grads else torch.no_grad()
    with ctx:
        out = fs2_oracle.forward(sd, cfg, batch, inference=meta["inference"], training=training, new_stats=new_stats)
        for k, want in gold.items():
            if not k.startswith("out."):
                continue
            got = out[k[4:]].detach().numpy()
            if want.dtype.kind in "biu":
                assert np.array_equal(got, want), k
            else:
                _close(got, want, 2e-5, f"{name}:{k}")
        if not meta["inference"]:
            losses = fs2_oracle.loss(out, batch, cfg, meta.get("epoch", 0))
            for k, want in gold.items():
                if k.startswith("loss."):
                    _close(float(losses[k[5:]]), float(want), 2e-5, f"{name}:{k}")
    if grads:
        losses["total"].backward()
        names = [str(n) for n in gold["grad.names"]]
        for n, norm, head in zip(names, gold["grad.norms"], gold["grad.heads"]):
            g = sd[n].grad
            assert g is not None, n
            g = g.double().flatten()
            assert abs(float(g.norm()) - norm) <= 1e-4 * max(norm, 1e-3), (n, float(g.norm()), norm)
        for k, want in gold.items():
            if k.startswith("bn."):
                _close(new_stats[k[3:]].numpy(), want, 1e-5, k)
