"""Device batch builder and prediction trimming against the restated reference collate
(fs2/dataset.py:257-293) and the callback's per-item trimming (fs2/prediction_writing_callback.py:255-262)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _items(B, learn_alignment, with_mel=True, seed=0):
    g = np.random.default_rng(seed)
    items = []
    for b in range(B):
        T = int(g.integers(5, 40))
        dur = g.integers(1, 6, size=T)
        F = int(dur.sum())
        it = {
            "text": torch.from_numpy(g.integers(1, 60, size=T).astype(np.int64)),
            "mel": torch.from_numpy(g.standard_normal((F, 80)).astype(np.float32)) if with_mel else None,
            "duration": (torch.from_numpy(g.random((F, T)).astype(np.float32)) if learn_alignment
                         else torch.from_numpy(dur.astype(np.int64))),
            "pitch": g.standard_normal(F if learn_alignment else T).astype(np.float32),  # ndarray on purpose (:271-272)
            "energy": torch.from_numpy(g.standard_normal(F if learn_alignment else T).astype(np.float32)),
            "speaker_id": int(g.integers(0, 4)),
            "language_id": int(g.integers(0, 2)),
            "basename": f"utt{b}",
            "duration_control": 1.0,
            "mel_style_reference": None,
        }
        items.append(it)
    return items


@pytest.mark.parametrize("learn_alignment,with_mel,B", [(True, True, 7), (False, True, 4), (False, False, 3), (True, True, 1)])
def test_collate_to_device_equals_reference_collate(learn_alignment, with_mel, B):
    from fastspeech2_lightning_b200.fs2.batching import collate_to_device
    from oracle.intops import collate_method

    items = _items(B, learn_alignment, with_mel, seed=B)
    ref = collate_method([dict(i) for i in items], learn_alignment)
    got = collate_to_device(items, DEV, learn_alignment)
    torch.cuda.synchronize()
    for k, v in ref.items():
        if torch.is_tensor(v) and v.dim() > 0:
            assert got[k].is_cuda and got[k].dtype == v.dtype and tuple(got[k].shape) == tuple(v.shape), (k, got[k].dtype, v.dtype, got[k].shape, v.shape)
            assert torch.equal(got[k].cpu(), v), k
        elif torch.is_tensor(v) or isinstance(v, int):
            assert int(got[k]) == int(v), k
        else:
            assert got[k] == v, k


def test_trim_predictions_equals_per_item_slicing():
    from fastspeech2_lightning_b200.fs2.batching import trim_predictions

    g = torch.Generator().manual_seed(1)
    B, F, C = 6, 211, 80
    mel = torch.randn(B, F, C, generator=g).to(DEV)
    lens = torch.tensor([211, 1, 64, 65, 200, 33], device=DEV)
    out = trim_predictions({"postnet_output": mel, "tgt_lens": lens})
    for b in range(B):
        ref = mel[b][: int(lens[b])].cpu().transpose(0, 1)
        assert out[b].shape == ref.shape and torch.equal(out[b], ref), b


def test_bucketed_padding_keeps_lengths_and_valid_values_and_the_step_runs():
    """Opt-in `pad_multiple` padding (fs2.batching): shapes are rounded up, lengths / valid values unchanged, padding is zero,
    and the training step accepts the batch; `collate_to_device(pad_multiple=...)` pads to the same shapes."""
    from fastspeech2_lightning_b200 import synthetic
    from fastspeech2_lightning_b200.fs2.batching import collate_to_device, pad_batch_to_multiple

    b = synthetic.make_batch(3, (20, 27), seed=12, learn_alignment=True)
    T, F = int(b["max_src_len"]), int(b["max_mel_len"])
    p = pad_batch_to_multiple(b, (8, 32))
    T2, F2 = int(p["max_src_len"]), int(p["max_mel_len"])
    assert (T2, F2) == ((T + 7) // 8 * 8, (F + 31) // 32 * 32) and (T2, F2) != (T, F)
    assert p["mel"].shape == (3, F2, 80) and p["duration"].shape == (3, F2, T2) and p["text"].shape == (3, T2) and p["pitch"].shape == (3, F2)
    assert torch.equal(p["mel"][:, :F], b["mel"]) and float(p["mel"][:, F:].abs().sum()) == 0.0
    assert torch.equal(p["duration"][:, :F, :T], b["duration"]) and torch.equal(p["src_lens"], b["src_lens"]) and torch.equal(p["mel_lens"], b["mel_lens"])
    assert pad_batch_to_multiple(p, (8, 32)) is p       # already on the grid
    items = _items(3, True, with_mel=True, seed=3)
    exact = collate_to_device(items, DEV, True)
    padded = collate_to_device(items, DEV, True, pad_multiple=(8, 32))
    Te, Fe = int(exact["max_src_len"]), int(exact["max_mel_len"])
    assert int(padded["max_src_len"]) == (Te + 7) // 8 * 8 and int(padded["max_mel_len"]) == (Fe + 31) // 32 * 32
    for k, v in exact.items():
        if torch.is_tensor(v) and v.dim() >= 2:
            w = padded[k]
            sl = tuple(slice(0, n) for n in v.shape)
            assert torch.equal(w[sl], v), k
            rest = w.clone()
            rest[sl] = 0
            assert float(rest.float().abs().sum()) == 0.0, k   # everything outside is zero
