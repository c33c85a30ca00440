"""Collate → device batch, and prediction trimming — the two host-side brackets of the hot path
(reference fs2/dataset.py:257-293 `FastSpeech2DataModule.collate_method`; fs2/prediction_writing_callback.py:255-262).

The reference pads every key on the host (`pad_sequence`, and a per-item loop that writes each [F_i,T_i] attention
prior into a zeroed [B,Fmax,Tmax] tensor) and the trainer then moves ~10 tensors to the GPU one by one.  Here the host
only copies the VALID values of every item back to back into ONE pinned staging buffer (plus the offset / shape
tables), does ONE host→device copy, and `fs2k_unpack_ragged` writes the padded tensors — zeros included — on the
device.  Same keys, dtypes, shapes and padding values as the reference's collate.

`trim_predictions` is the read side: the callback's per-item `data[:len].cpu().transpose(0, 1)` becomes one
`fs2k_trim_transpose` launch that packs every utterance's valid frames as [n_mels, T_b], one device→host copy, and
views into the pinned result.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .._lib import check, lib

_WORD = 4


def _flatten(item: dict) -> dict:
    """Nested dicts → one level (everyvoice.utils._flatten is un-vendored; leaf keys are kept as they are)."""
    out = {}
    for k, v in item.items():
        if isinstance(v, dict):
            out.update(_flatten(v))
        else:
            out[k] = v
    return out


def _as_numpy(x):
    if torch.is_tensor(x):
        return x.detach().cpu().numpy()
    return x


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m if m and m > 1 else n


def collate_to_device(data, device, learn_alignment: bool = True, pad_multiple=None) -> dict:
    """`collate_method(data, learn_alignment)` with the result on `device` (see the module docstring).

    `pad_multiple=(text_multiple, mel_multiple)` (opt-in, not reference behaviour): pad to the next multiple instead of to
    the batch maximum (dataset.py:257-293), so that a stream of batches falls into a few (T, F) shapes and
    `FastSpeech2.optimization_step` replays captured graphs instead of running every new shape eagerly.  Padding is live
    in this model (BatchNorm batch statistics, convolution edges and the full-rectangle loss means see it), so results
    differ from the exact-shape batch exactly as they would if the batch contained one longer utterance; lengths, masks
    and every valid position of the inputs are unchanged.

    Tensor / ndarray valued keys are padded on the device; int keys become int32 tensors; everything else (strings,
    None, floats) stays a Python list, as in the reference.  `max_src_len` / `max_mel_len` are 0-d int32 host tensors
    (the reference's `max(text_lens)`), `mel_lens` is None and `max_mel_len` 1_000_000 when the items carry no mel."""
    data = [_flatten(x) for x in data]
    keys = list(data[0])
    cols = {k: [d[k] for d in data] for k in keys}
    B = len(data)
    text_lens = np.array([int(t.shape[0]) for t in cols["text"]], dtype=np.int32)
    max_text = int(text_lens.max())
    has_mel = cols.get("mel", [None])[0] is not None
    mel_lens = np.array([int(m.shape[0]) for m in cols["mel"]], dtype=np.int32) if has_mel else None
    max_mel = int(mel_lens.max()) if has_mel else 1_000_000
    exact_text, exact_mel = max_text, max_mel
    if pad_multiple is not None:
        max_text = _round_up(max_text, int(pad_multiple[0]))
        if has_mel:
            max_mel = _round_up(max_mel, int(pad_multiple[1]))

    # ---- plan: every array-valued key becomes (items as contiguous 4-byte-word arrays, rows, cols, out shape / dtype)
    plans, small = [], {}
    n_words = 0
    for k in keys:
        first = cols[k][0]
        if isinstance(first, np.ndarray) or torch.is_tensor(first):
            items = [np.ascontiguousarray(_as_numpy(x)) for x in cols[k]]
            two_sided = k == "duration" and learn_alignment
            if two_sided:  # torch.zeros(B, max_mel, max_text) receives the values: float32 (dataset.py:276-281)
                items = [np.ascontiguousarray(x, dtype=np.float32) for x in items]
            dt = items[0].dtype
            if dt.itemsize % _WORD:
                raise TypeError(f"collate_to_device: key {k!r} has dtype {dt}; only 4- and 8-byte element types are packed")
            wpe = dt.itemsize // _WORD  # words per element
            trailing = int(np.prod(items[0].shape[1:])) if items[0].ndim > 1 else 1
            if two_sided:
                rows = [x.shape[0] for x in items]
                cws = [x.shape[1] * wpe for x in items]
                out_shape, rmax, cmax = (B, max_mel, max_text), max_mel, max_text * wpe
            else:  # pad_sequence(batch_first=True): ragged first dim, equal trailing dims
                rows = [x.shape[0] for x in items]
                cws = [trailing * wpe] * B
                rmax = max(rows)
                if pad_multiple is not None:  # frame-level keys follow the padded mel length, token-level keys the padded text length
                    rmax = max_mel if (has_mel and rmax == exact_mel) else (max_text if rmax == exact_text else rmax)
                out_shape, cmax = (B, rmax) + tuple(items[0].shape[1:]), trailing * wpe
            offs = []
            for x in items:
                offs.append(n_words)
                n_words += x.size * wpe
            plans.append((k, items, rows, cws, offs, out_shape, torch.from_numpy(np.empty(0, dtype=dt)).dtype, rmax, cmax))
        elif isinstance(first, (int, np.integer)) and not isinstance(first, bool):
            small[k] = np.array(cols[k], dtype=np.int32)

    # ---- one pinned staging buffer: [packed words | per-plan tables (offsets int64, rows int32, cols int32) | small int keys]
    table_words = sum(2 * B + B + B for _ in plans)
    small_words = sum(len(v) for v in small.values()) + B + (B if has_mel else 0)
    n_words_al = (n_words + 1) // 2 * 2  # keep the int64 tables 8-byte aligned
    stage = torch.empty((n_words_al + table_words + small_words) * _WORD, dtype=torch.uint8).pin_memory()
    st = stage.numpy()
    words = st.view(np.uint32)
    for k, items, rows, cws, offs, *_ in plans:
        for x, o in zip(items, offs):
            words[o: o + x.size * (x.dtype.itemsize // _WORD)] = x.reshape(-1).view(np.uint32)
    pos = n_words_al
    table_pos = []
    for k, items, rows, cws, offs, *_ in plans:
        words[pos: pos + 2 * B].view(np.int64)[:] = np.asarray(offs, dtype=np.int64)
        words[pos + 2 * B: pos + 3 * B].view(np.int32)[:] = np.asarray(rows, dtype=np.int32)
        words[pos + 3 * B: pos + 4 * B].view(np.int32)[:] = np.asarray(cws, dtype=np.int32)
        table_pos.append(pos)
        pos += 4 * B
    small_pos = {}
    for k, v in list(small.items()) + [("src_lens", text_lens)] + ([("mel_lens", mel_lens)] if has_mel else []):
        words[pos: pos + len(v)].view(np.int32)[:] = v
        small_pos[k] = (pos, len(v))
        pos += len(v)

    dev_stage = stage.to(device, non_blocking=True)  # THE host→device copy of this batch
    dev_words = dev_stage.view(torch.int32)
    out = dict(cols)
    stream = torch.cuda.current_stream(device).cuda_stream
    base = dev_stage.data_ptr()
    for (k, items, rows, cws, offs, out_shape, tdtype, rmax, cmax), tp in zip(plans, table_pos):
        t = torch.empty(out_shape, dtype=tdtype, device=device)
        check(lib().fs2k_unpack_ragged(base, base + tp * _WORD, base + (tp + 2 * B) * _WORD, base + (tp + 3 * B) * _WORD, B, rmax, cmax,
                                       t.data_ptr(), stream), "fs2k_unpack_ragged")
        ops._count()
        out[k] = t
    for k, (p, n) in small_pos.items():
        out[k] = dev_words[p: p + n]
    out["max_src_len"] = torch.tensor(max_text, dtype=torch.int32)
    if not has_mel:
        out["mel_lens"] = None
    out["max_mel_len"] = torch.tensor(max_mel, dtype=torch.int32) if has_mel else max_mel
    out["_staging"] = dev_stage  # keeps the small views' storage alive with the batch
    return out


def pad_batch_to_multiple(batch: dict, pad_multiple) -> dict:
    """The opt-in bucketed padding of `collate_to_device(pad_multiple=...)` for an already collated batch dict: every tensor
    dimension that is the batch's `max_src_len` (`max_mel_len`) is zero-padded to the next multiple of `pad_multiple[0]`
    (`[1]`), and the two maxima are updated.  Lengths and valid values are untouched.  Not reference behaviour — see
    `collate_to_device`."""
    T, F = int(batch["max_src_len"]), int(batch["max_mel_len"])
    has_mel = batch.get("mel_lens") is not None and F < 1_000_000
    T2 = _round_up(T, int(pad_multiple[0]))
    F2 = _round_up(F, int(pad_multiple[1])) if has_mel else F
    if (T2, F2) == (T, F):
        return batch
    if has_mel and T == F:
        raise ValueError("pad_batch_to_multiple cannot tell token-level from frame-level keys when max_src_len == max_mel_len")
    out = dict(batch)
    for k, v in batch.items():
        if not torch.is_tensor(v) or v.dim() < 2:
            continue
        # dimension 1 is the ragged one (pad_sequence); only the attention prior `duration` [B,F,T] is ragged in two
        pads = []
        for d in range(v.dim() - 1, 0, -1):  # F.pad takes the last dimension first
            n = v.shape[d]
            ragged = d == 1 or (k == "duration" and v.dim() == 3)
            target = n if not ragged else (F2 if (has_mel and n == F) else (T2 if n == T else n))
            pads += [0, target - n]
        if any(pads):
            out[k] = torch.nn.functional.pad(v, pads)
    out["max_src_len"] = torch.tensor(T2, dtype=torch.int32)
    if has_mel:
        out["max_mel_len"] = torch.tensor(F2, dtype=torch.int32)
    return out


_pinned = {}  # device index → reusable pinned result buffer (cudaHostAlloc per call would cost more than the copy)


def _pinned_buffer(device, n: int) -> torch.Tensor:
    buf = _pinned.get(device.index)
    if buf is None or buf.numel() < n:
        buf = _pinned[device.index] = torch.empty(max(n, 1 << 20), dtype=torch.float32).pin_memory()
    return buf


def trim_predictions(outputs: dict, output_key: str = "postnet_output", reuse_buffer: bool = False) -> list[torch.Tensor]:
    """`[data[:tgt_len].cpu().transpose(0, 1) for data in outputs[output_key]]` (prediction_writing_callback.py:255-262)
    as one device launch + one device→host copy.  Returns CPU tensors [n_mels, T_b] (views of one pinned buffer; with
    `reuse_buffer` the buffer is shared between calls — consume the result before the next call, as the callback does)."""
    mel = outputs[output_key]
    lens = outputs["tgt_lens"]
    assert mel is not None and lens is not None
    B, F, C = mel.shape
    lens_host = lens.to("cpu", torch.int64).clamp_(max=F)  # the one host read the file names/sizes need anyway
    offs = torch.zeros(B + 1, dtype=torch.int64)
    torch.cumsum(lens_host * C, 0, out=offs[1:])
    total = int(offs[-1])
    dev_offs = offs[:B].to(mel.device, non_blocking=True)
    packed = torch.empty(max(total, 1), dtype=torch.float32, device=mel.device)
    check(lib().fs2k_trim_transpose(mel.contiguous().data_ptr(), lens.to(torch.int32).contiguous().data_ptr(), dev_offs.data_ptr(), B, F, C,
                                    packed.data_ptr(), torch.cuda.current_stream(mel.device).cuda_stream), "fs2k_trim_transpose")
    ops._count()
    host = _pinned_buffer(mel.device, total)[: max(total, 1)] if reuse_buffer else torch.empty(max(total, 1), dtype=torch.float32).pin_memory()
    host.copy_(packed, non_blocking=False)
    return [host[int(offs[b]): int(offs[b + 1])].view(C, int(lens_host[b])) for b in range(B)]
