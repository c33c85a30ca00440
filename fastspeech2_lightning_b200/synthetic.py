"""Seeded synthetic weights and batches (there is no dataset and no checkpoint).

Everything is drawn from `numpy.random.Generator(PCG64)` streams — stable across
numpy/torch versions — so the golden fixtures under `tests/golden/` (made from the
reference in the build container) and the tensors rebuilt on the GPU box agree
bit for bit.  The batch dict has the layout of the reference's collate function
(`fs2/dataset.py:257-293`).
"""
from __future__ import annotations

import zlib
from typing import Optional

import numpy as np
import torch


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(key.encode())])


def synth_tensor(key: str, shape, seed: int) -> Optional[torch.Tensor]:
    """Deterministic fp32 value for state-dict entry `key` (None = leave untouched)."""
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    g = _rng(seed, key)
    if leaf in ("inv_freq", "pitch_bins", "energy_bins"):
        return None
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    n = g.standard_normal(shape)
    if leaf == "running_mean":
        v = 0.1 * n
    elif leaf == "running_var":
        v = 1.0 + 0.25 * np.abs(n)
    elif leaf in ("bias", "in_proj_bias"):
        v = 0.05 * n
    elif leaf == "gst_embs":
        v = n
    elif leaf in ("weight", "in_proj_weight", "weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
        if len(shape) == 1:  # LayerNorm / BatchNorm scale
            v = 1.0 + 0.1 * n
        elif key.endswith("_embedding.weight"):  # pitch/energy/speaker/language tables
            v = 0.3 * n
        elif key == "text_input_layer.weight" and shape[-1] != 39:  # symbol table (not the pfs Linear)
            v = n
        else:
            fan_in = int(np.prod(shape[1:]))
            v = n / np.sqrt(fan_in)
    else:
        v = 0.05 * n
    return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(torch.float32)


@torch.no_grad()
def fill_weights_(module: torch.nn.Module, seed: int = 0, pad_row_zero: bool = True) -> None:
    """Overwrite every parameter/buffer of `module` with the seeded synthetic value."""
    sd = module.state_dict()
    for k, t in sd.items():
        v = synth_tensor(k, t.shape, seed)
        if v is None:
            continue
        if k == "text_input_layer.weight" and pad_row_zero and v.dim() == 2 and v.shape[1] != 39:
            v[0].zero_()  # padding_idx row (model.py:83-89)
        t.copy_(v.to(t.dtype))


def beta_binomial_prior(n_text: int, n_mel: int, scaling: float = 1.0) -> torch.Tensor:
    """Beta-binomial alignment prior [n_mel, n_text] (the upstream preprocessor's
    attention prior; Badlani et al. 2021 §3.2): row i ~ BetaBinom(n_text-1, s·i, s·(n_mel+1-i))."""
    k = torch.arange(n_text, dtype=torch.float64)[None, :]
    i = torch.arange(1, n_mel + 1, dtype=torch.float64)[:, None]
    a, b, n = scaling * i, scaling * (n_mel + 1 - i), float(n_text - 1)
    lg = torch.lgamma
    logc = lg(torch.tensor(n + 1.0, dtype=torch.float64)) - lg(k + 1) - lg(n - k + 1)
    logp = logc + lg(k + a) + lg(n - k + b) - lg(n + a + b) - (lg(a) + lg(b) - lg(a + b))
    return torch.exp(logp).to(torch.float32)


def make_batch(
    batch_size: int,
    src_len_range: tuple[int, int],
    *,
    seed: int = 1234,
    n_symbols: int = 27,
    n_mels: int = 80,
    dur_range: tuple[int, int] = (3, 9),
    learn_alignment: bool = True,
    inference: bool = False,
    teacher_forced: bool = True,
    n_speakers: int = 1,
    n_languages: int = 1,
    first_is_longest: bool = True,
    src_lens=None,
    style_frames: int = 0,
    device: str | torch.device = "cpu",
) -> dict:
    """A batch dict as `FastSpeech2DataModule.collate_method` would build it.

    * training / learned alignment: `duration` is the attention prior f32 [B,F,T],
      `pitch`/`energy` frame level f32 [B,F], `mel` f32 [B,F,n_mels].
    * `learn_alignment=False`: `duration` int32 [B,T], pitch/energy phone level [B,T].
    * `inference=True`: `mel=None`, pitch=energy=None; with `teacher_forced` the integer
      durations and `mel_lens` are supplied (model.py:162-165), otherwise `mel_lens=None`,
      `max_mel_len=1_000_000` (dataset.py:263-268).
    * `src_lens=[...]` fixes the phone count of every utterance; `style_frames > 0` adds a GST style reference mel.
    """
    g = _rng(seed, "batch")
    lo, hi = src_len_range
    if src_lens is not None:  # explicit phone counts (length-sorted batches of a corpus: bench synth_c4)
        src_lens = np.asarray(src_lens, dtype=np.int64)
        batch_size = len(src_lens)
    else:
        src_lens = g.integers(lo, hi + 1, size=batch_size)
        if first_is_longest:
            src_lens[0] = hi
    T = int(src_lens.max())
    text = np.zeros((batch_size, T), dtype=np.int32)
    dur = np.zeros((batch_size, T), dtype=np.int32)
    for b in range(batch_size):
        text[b, : src_lens[b]] = g.integers(1, n_symbols, size=src_lens[b])
        dur[b, : src_lens[b]] = g.integers(dur_range[0], dur_range[1] + 1, size=src_lens[b])
    mel_lens = dur.sum(1).astype(np.int32)
    F = int(mel_lens.max())
    batch: dict = {
        "text": torch.from_numpy(text),
        "src_lens": torch.from_numpy(src_lens.astype(np.int32)),
        "max_src_len": torch.tensor(T, dtype=torch.int32),
        "speaker_id": torch.from_numpy(g.integers(0, n_speakers, size=batch_size).astype(np.int32)),
        "language_id": torch.from_numpy(g.integers(0, n_languages, size=batch_size).astype(np.int32)),
        "duration_control": [1.0] * batch_size,
        "mel_style_reference": [None] * batch_size,
        "pfs": None,
        "basename": [f"synthetic-{seed}-{b}" for b in range(batch_size)],
        "speaker": ["default"] * batch_size,
        "language": ["default"] * batch_size,
        "raw_text": [""] * batch_size,
    }
    if style_frames:  # GST style reference mel [B, style_frames, n_mels] (fs2/model.py:196-199)
        batch["mel_style_reference"] = torch.from_numpy(g.standard_normal((batch_size, style_frames, n_mels)).astype(np.float32))
    frame_valid = np.arange(F)[None, :] < mel_lens[:, None]
    if inference:
        batch["mel"] = None
        batch["energy"] = None
        batch["pitch"] = None
        if teacher_forced:
            batch["duration"] = torch.from_numpy(dur)
            batch["mel_lens"] = torch.from_numpy(mel_lens)
            batch["max_mel_len"] = torch.tensor(F, dtype=torch.int32)
        else:
            batch["duration"] = [None] * batch_size
            batch["mel_lens"] = None
            batch["max_mel_len"] = 1_000_000
    else:
        mel = g.standard_normal((batch_size, F, n_mels)).astype(np.float32) * frame_valid[:, :, None]
        batch["mel"] = torch.from_numpy(mel.astype(np.float32))
        batch["mel_lens"] = torch.from_numpy(mel_lens)
        batch["max_mel_len"] = torch.tensor(F, dtype=torch.int32)
        if learn_alignment:
            prior = torch.zeros(batch_size, F, T)
            for b in range(batch_size):
                prior[b, : mel_lens[b], : src_lens[b]] = beta_binomial_prior(int(src_lens[b]), int(mel_lens[b]))
            batch["duration"] = prior
            pitch = g.standard_normal((batch_size, F)).astype(np.float32) * frame_valid
            energy = g.standard_normal((batch_size, F)).astype(np.float32) * frame_valid
            # unvoiced frames are exact zeros in real pitch tracks; average_variance skips them
            pitch = pitch * (g.random((batch_size, F)) > 0.2)
        else:
            batch["duration"] = torch.from_numpy(dur)
            src_valid = np.arange(T)[None, :] < src_lens[:, None]
            pitch = g.standard_normal((batch_size, T)).astype(np.float32) * src_valid
            energy = g.standard_normal((batch_size, T)).astype(np.float32) * src_valid
        batch["pitch"] = torch.from_numpy(pitch.astype(np.float32))
        batch["energy"] = torch.from_numpy(energy.astype(np.float32))
    return batch_to(batch, device)


def batch_to(batch: dict, device, non_blocking: bool = False) -> dict:
    out = {}
    for k, v in batch.items():
        # 0-d shape scalars (max_src_len / max_mel_len) stay on the host: the model only reads them as ints
        out[k] = v.to(device, non_blocking=non_blocking) if (torch.is_tensor(v) and v.dim() > 0) else v
    return out


DEFAULT_STATS = {
    "pitch": dict(min=0.0, max=1.0, std=1.0, mean=0.0, norm_min=-3.0, norm_max=3.0),
    "energy": dict(min=0.0, max=1.0, std=1.0, mean=0.0, norm_min=-3.0, norm_max=3.0),
}
