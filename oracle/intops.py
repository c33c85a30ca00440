"""ORACLE (test infrastructure, not product code) — integer/index arithmetic of the
hot path restated in numpy / plain Python from the reference.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
arm may import this.  Parity status: PINNED — every function here is checked against
the reference's own code run in the build container (`tests/golden/make_golden.py`
→ `tests/golden/kat_*.npz`, `tests/test_oracle.py`).  The reference's tests hold no
golden vectors for this path, so the fixtures are outputs of the
unmodified reference functions.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent


# ---------------------------------------------------------------------------
# monotonic alignment search — reference fs2/attn/alignment.py:48-74 (mas_width1)
# ---------------------------------------------------------------------------
def mas_width1_py(log_attn_map: np.ndarray) -> np.ndarray:
    """Pure-Python restatement (small cases only).  `log_attn_map` is mel × text."""
    log_p = np.array(log_attn_map, dtype=np.float32, copy=True)
    n_mel, n_text = log_p.shape
    neg_inf = np.float32(-np.inf)
    log_p[0, 1:] = neg_inf  # :54
    for i in range(1, n_mel):  # :55-60
        prev1 = neg_inf
        for j in range(n_text):
            prev2 = log_p[i - 1, j]
            log_p[i, j] = np.float32(log_p[i, j] + max(prev1, prev2))
            prev1 = prev2
    opt = np.zeros_like(log_p)
    j = n_text - 1
    for i in range(n_mel - 1, 0, -1):  # :66-72
        opt[i, j] = 1.0
        # NB: j == 0 reads log_p[i-1, -1] (numpy wrap-around) exactly like the reference
        if log_p[i - 1, j - 1] >= log_p[i - 1, j]:
            j -= 1
            if j == 0:
                opt[1:i, j] = 1.0
                break
    opt[0, j] = 1.0  # :73
    return opt


_lib = None


def _clib() -> ctypes.CDLL:
    """Builds (gcc) and loads oracle/_build/libfs2oracle.so — the C restatement in oracle/mas.c."""
    global _lib
    if _lib is not None:
        return _lib
    so = _HERE / "_build" / "libfs2oracle.so"
    src = _HERE / "mas.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        so.parent.mkdir(exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-fno-fast-math", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", str(so), str(src)]
        )
    _lib = ctypes.CDLL(str(so))
    _lib.oracle_mas_width1.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_long]
    _lib.oracle_b_mas.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return _lib


def mas_width1(log_attn_map: np.ndarray) -> np.ndarray:
    """C restatement of mas_width1 (any size)."""
    x = np.ascontiguousarray(log_attn_map, dtype=np.float32)
    n_mel, n_text = x.shape
    out = np.zeros_like(x)
    scratch = np.empty_like(x)
    _clib().oracle_mas_width1(x.ctypes.data, out.ctypes.data, scratch.ctypes.data, n_mel, n_text, n_text, n_text)
    return out


def b_mas(b_log_attn_map: np.ndarray, in_lens, out_lens, threads: int = 0) -> np.ndarray:
    """Batched MAS on [B,1,F,T] — semantics of reference b_mas (alignment.py:77-85) and of the
    loop in VarianceAdaptor.binarize_attention (variance_adaptor.py:173-179): per item, MAS on
    [:out_len, :in_len], zeros elsewhere."""
    x = np.ascontiguousarray(b_log_attn_map, dtype=np.float32)
    B, one, F, T = x.shape
    assert one == 1
    il = np.ascontiguousarray(in_lens, dtype=np.int32)
    ol = np.ascontiguousarray(out_lens, dtype=np.int32)
    out = np.zeros_like(x)
    _clib().oracle_b_mas(x.ctypes.data, out.ctypes.data, il.ctypes.data, ol.ctypes.data, B, F, T, threads)
    return out


# ---------------------------------------------------------------------------
# LengthRegulator — reference fs2/variance_adaptor.py:65-81
# ---------------------------------------------------------------------------
def length_regulator(x: np.ndarray, durations: np.ndarray, max_length: int):
    """out[b] = repeat_interleave(x[b], dur[b]) zero-padded to width min(max_b Σdur, max_length);
    mask[b,f] = f < Σdur[b] (the un-truncated length, :74-77).  Also returns the implied gather
    index idx[b,f] (−1 on padding) that the CUDA kernel must reproduce bit-exactly."""
    B, T = durations.shape
    lengths = durations.astype(np.int64).sum(1)
    width = int(min(lengths.max(), int(max_length)))
    out = np.zeros((B, width) + x.shape[2:], dtype=x.dtype)
    idx = np.full((B, width), -1, dtype=np.int32)
    for b in range(B):
        rep = np.repeat(np.arange(T), durations[b])[:width]
        idx[b, : len(rep)] = rep
        out[b, : len(rep)] = x[b, rep]
    mask = np.arange(width)[None, :] < lengths[:, None]
    return out, mask, idx


# ---------------------------------------------------------------------------
# bucketize — torch.bucketize(v, bins) with right=False (variance_adaptor.py:197-204)
# ---------------------------------------------------------------------------
def bucketize(v: np.ndarray, bins: np.ndarray) -> np.ndarray:
    """id = #{bins < v} (lower bound); NaN sorts after everything → len(bins)."""
    v = np.asarray(v, dtype=np.float32)
    ids = np.searchsorted(np.asarray(bins, dtype=np.float32), v, side="left").astype(np.int64)
    ids[np.isnan(v)] = len(bins)
    return ids


# ---------------------------------------------------------------------------
# average_variance — reference fs2/variance_adaptor.py:207-222
# ---------------------------------------------------------------------------
def average_variance(var: np.ndarray, durs: np.ndarray) -> np.ndarray:
    """Per phone: (C[end]-C[start]) / (N[end]-N[start]) with C the fp32 running prefix sum of the
    frame values and N the prefix count of non-zero frames; 0 where the count is 0.  The prefix
    sums are accumulated sequentially in fp32 — the order torch's CPU cumsum uses."""
    var = np.asarray(var, dtype=np.float32)
    B, F = var.shape
    T = durs.shape[1]
    out = np.zeros((B, T), dtype=np.float32)
    for b in range(B):
        c = np.zeros(F + 1, dtype=np.float32)
        n = np.zeros(F + 1, dtype=np.int64)
        acc = np.float32(0.0)
        for f in range(F):
            acc = np.float32(acc + var[b, f])
            c[f + 1] = acc
            n[f + 1] = n[f] + (var[b, f] != 0.0)
        ends = np.cumsum(durs[b].astype(np.int64))
        starts = np.concatenate([[0], ends[:-1]])
        s = (c[ends] - c[starts]).astype(np.float32)
        k = (n[ends] - n[starts]).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            out[b] = np.where(k == 0.0, k, s / k)
    return out


# ---------------------------------------------------------------------------
# inference duration rounding — reference fs2/variance_adaptor.py:359-366
# ---------------------------------------------------------------------------
def round_durations(log_dur: np.ndarray, control: float = 1.0) -> np.ndarray:
    """clamp(round_half_even(exp(logd) - 1) * control, min=0) truncated to int32.  `exp` is torch's
    (the result at an exact .5 boundary depends on the last ulp of exp, so the oracle uses the same
    library routine as the reference)."""
    import torch

    x = torch.from_numpy(np.asarray(log_dur, dtype=np.float32))
    return torch.clamp(torch.round(torch.exp(x) - 1) * control, min=0).int().numpy()


def collate_method(data, learn_alignment=True):
    """Restatement of FastSpeech2DataModule.collate_method (fs2/dataset.py:257-293) for already-flat item dicts:
    `pad_sequence(batch_first=True, padding_value=0)` for tensor keys, the two-sided zero padding of the [F_i,T_i]
    attention prior into float32 [B,Fmax,Tmax] when `learn_alignment`, IntTensor for int keys, lens / max lens."""
    import numpy as np
    import torch
    from torch.nn.utils.rnn import pad_sequence

    data = {k: [dic[k] for dic in data] for k in data[0]}
    text_lens = torch.IntTensor([text.size(0) for text in data["text"]])
    max_text = max(text_lens)
    if data["mel"][0] is not None:
        mel_lens = torch.IntTensor([mel.size(0) for mel in data["mel"]])
        max_mel = max(mel_lens)
    else:
        mel_lens, max_mel = None, 1_000_000
    for key in data:
        if isinstance(data[key][0], np.ndarray):
            data[key] = [torch.tensor(x) for x in data[key]]
        if torch.is_tensor(data[key][0]):
            if key == "duration" and learn_alignment:
                padded = torch.zeros(len(text_lens), max_mel, max_text)
                for i, dur in enumerate(data[key]):
                    padded[i, : dur.size(0), : dur.size(1)] = dur
                data[key] = padded
            else:
                data[key] = pad_sequence(data[key], batch_first=True, padding_value=0)
        if isinstance(data[key][0], int):
            data[key] = torch.IntTensor(data[key])
    data["src_lens"], data["max_src_len"], data["mel_lens"], data["max_mel_len"] = text_lens, max_text, mel_lens, max_mel
    return data
