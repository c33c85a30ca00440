"""Tensor-level wrappers over the libfs2k C ABI.

PyTorch is used for device memory and streams only: every function takes CUDA tensors,
allocates its outputs with `torch.empty`, and enqueues hand-written sm_100a kernels on the
current stream.  Nothing here computes with torch ops; there is no CPU path.
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch

from ._lib import check, lib

ACT_NONE, ACT_RELU, ACT_SILU, ACT_TANH = 0, 1, 2, 3
_ACTS = {None: 0, "none": 0, "relu": 1, "silu": 2, "tanh": 3}

# count of kernel launches issued through this module (bench.py reports it as gpu_launches)
launch_count = 0


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream() -> int:
    """The current stream's handle.  Asked once per launch (≈ 850 times per training step), so it goes through the two C-level
    calls torch's own extensions use; `torch.cuda.current_stream()` builds a Stream object and re-checks the device and
    the environment on every call (6 ms of host time per eager step)."""
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f32(t: torch.Tensor, name: str = "tensor") -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _i32(t: torch.Tensor, name: str = "tensor") -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU path)")
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t if t.is_contiguous() else t.contiguous()


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


# ---------------------------------------------------------------------------------------------
# monotonic alignment search
# ---------------------------------------------------------------------------------------------
def mas(attn: torch.Tensor, in_lens: torch.Tensor, out_lens: torch.Tensor, take_log: bool = False, dense: bool = True):
    """attn [B,1,F,T] or [B,F,T] (log-probs, or probs with take_log) → (path [B,F] i32, durations [B,T] i32,
    hard [B,1,F,T] f32 or None)."""
    a = _f32(attn, "attn")
    if a.dim() == 4:
        assert a.shape[1] == 1
        B, _, F, T = a.shape
    else:
        B, F, T = a.shape
    il, ol = _i32(in_lens, "in_lens"), _i32(out_lens, "out_lens")
    path = torch.empty((B, F), dtype=torch.int32, device=a.device)
    dur = torch.empty((B, T), dtype=torch.int32, device=a.device)
    hard = torch.empty((B, 1, F, T), dtype=torch.float32, device=a.device) if dense else None
    ws_bytes = lib().fs2k_mas_workspace_bytes(B, F, T)
    ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=a.device)
    check(lib().fs2k_mas_fwd(_p(a), int(take_log), _p(il), _p(ol), B, F, T, _p(path), _p(dur), _p(hard), _p(ws), ws_bytes, _stream()), "fs2k_mas_fwd")
    _count((2 if dense else 1) + (1 if take_log else 0))
    return path, dur, hard


# ---------------------------------------------------------------------------------------------
# length regulator
# ---------------------------------------------------------------------------------------------
def lr_scan(durations: torch.Tensor):
    d = _i32(durations, "durations")
    B, T = d.shape
    cum = torch.empty_like(d)
    total = torch.empty((B,), dtype=torch.int32, device=d.device)
    check(lib().fs2k_lr_scan(_p(d), B, T, _p(cum), _p(total), _stream()), "fs2k_lr_scan")
    _count()
    return cum, total


def lr_gather(x, cum, total, f_out: int, inv_freq: Optional[torch.Tensor] = None, want_out=True, want_idx=False):
    x = _f32(x, "x")
    B, T, D = x.shape
    dev = x.device
    out = torch.empty((B, f_out, D), dtype=torch.float32, device=dev) if want_out else None
    out_pos = torch.empty((B, f_out, D), dtype=torch.float32, device=dev) if inv_freq is not None else None
    mask = torch.empty((B, f_out), dtype=torch.bool, device=dev)
    idx = torch.empty((B, f_out), dtype=torch.int32, device=dev) if want_idx else None
    check(lib().fs2k_lr_gather(_p(x), _p(cum), _p(total), B, T, D, f_out, _p(out), _p(out_pos), _p(inv_freq), _p(mask), _p(idx), _stream()), "fs2k_lr_gather")
    _count()
    return out, out_pos, mask, idx


# ---------------------------------------------------------------------------------------------
# variance adaptor pieces
# ---------------------------------------------------------------------------------------------
def bucketize_embed_add(v, bins, table, x, scale: float = 1.0, want_ids=True, want_scaled=False):
    v, bins, table, x = _f32(v, "v"), _f32(bins, "bins"), _f32(table, "table"), _f32(x, "x")
    N, D = v.numel(), x.shape[-1]
    assert x.numel() == N * D and table.shape[1] == D and table.shape[0] >= bins.numel() + 1
    y = torch.empty_like(x)
    ids = torch.empty(v.shape, dtype=torch.int64, device=v.device) if want_ids else None
    vs = torch.empty_like(v) if want_scaled else None
    check(lib().fs2k_bucketize_embed_add(_p(v), float(scale), _p(vs), _p(bins), bins.numel(), _p(table), _p(x), _p(y), _p(ids), N, D, _stream()), "fs2k_bucketize_embed_add")
    _count()
    return y, ids, vs


def bucketize(v, bins):
    v, bins = _f32(v, "v"), _f32(bins, "bins")
    ids = torch.empty(v.shape, dtype=torch.int64, device=v.device)
    check(lib().fs2k_bucketize(_p(v), _p(bins), bins.numel(), _p(ids), v.numel(), _stream()), "fs2k_bucketize")
    _count()
    return ids


def average_variance(var, cum):
    var = _f32(var, "var")
    B, F = var.shape
    T = cum.shape[1]
    out = torch.empty((B, T), dtype=torch.float32, device=var.device)
    check(lib().fs2k_average_variance(_p(var), _p(cum), B, F, T, _p(out), _stream()), "fs2k_average_variance")
    _count()
    return out


def round_durations(log_dur, control: float = 1.0):
    ld = _f32(log_dur, "log_dur")
    dur = torch.empty(ld.shape, dtype=torch.int32, device=ld.device)
    check(lib().fs2k_round_durations(_p(ld), float(control), ld.numel(), _p(dur), _stream()), "fs2k_round_durations")
    _count()
    return dur


def embed_posenc(text, table, inv_freq, lens, want_emb=True):
    text, table, inv_freq, lens = _i32(text, "text"), _f32(table, "table"), _f32(inv_freq, "inv_freq"), _i32(lens, "lens")
    B, T = text.shape
    D = table.shape[1]
    emb = torch.empty((B, T, D), dtype=torch.float32, device=table.device) if want_emb else None
    x = torch.empty((B, T, D), dtype=torch.float32, device=table.device)
    check(lib().fs2k_embed_posenc(_p(text), _p(table), table.shape[0], _p(inv_freq), _p(lens), B, T, D, _p(emb), _p(x), None, _stream()), "fs2k_embed_posenc")
    _count()
    return emb, x


def add_posenc(x, inv_freq, lens):
    x, inv_freq, lens = _f32(x, "x"), _f32(inv_freq, "inv_freq"), _i32(lens, "lens")
    B, L, D = x.shape
    y = torch.empty_like(x)
    check(lib().fs2k_add_posenc(_p(x), _p(inv_freq), _p(lens), B, L, D, _p(y), _stream()), "fs2k_add_posenc")
    _count()
    return y


def add_rows(x, rows_and_ids):
    """x [B,L,D] + Σ rows[ids[b]] for up to three (rows [R,D], ids [B] int32 or None) pairs."""
    x = _f32(x, "x")
    B, L, D = x.shape
    y = torch.empty_like(x)
    args = []
    keep = []
    for rows, ids in list(rows_and_ids) + [(None, None)] * (3 - len(rows_and_ids)):
        if rows is not None:
            rows = _f32(rows, "rows")
            ids = _i32(ids, "ids") if ids is not None else None
            keep += [rows, ids]
        args += [_p(rows), _p(ids)]
    check(lib().fs2k_add_rows(_p(x), _p(y), B, L, D, *args, _stream()), "fs2k_add_rows")
    _count()
    return y


def lens_mask(lens, max_len: int):
    lens = _i32(lens, "lens")
    B = lens.shape[0]
    mask = torch.empty((B, int(max_len)), dtype=torch.bool, device=lens.device)
    check(lib().fs2k_lens_mask(_p(lens), B, int(max_len), _p(mask), _stream()), "fs2k_lens_mask")
    _count()
    return mask


def mask_lens(mask):
    assert mask.is_cuda and mask.dtype in (torch.bool, torch.uint8)
    m = mask.contiguous()
    B, L = m.shape
    lens = torch.empty((B,), dtype=torch.int32, device=m.device)
    check(lib().fs2k_mask_lens(_p(m), B, L, _p(lens), _stream()), "fs2k_mask_lens")
    _count()
    return lens


# ---------------------------------------------------------------------------------------------
# normalisation
# ---------------------------------------------------------------------------------------------
def layernorm(x, gamma, beta, eps: float = 1e-5, save_stats: bool = False, dropout_p: float = 0.0, seed: int = 0):
    x = _f32(x, "x")
    D = x.shape[-1]
    M = x.numel() // D
    y = torch.empty_like(x)
    mean = torch.empty((M,), dtype=torch.float32, device=x.device) if save_stats else None
    rstd = torch.empty((M,), dtype=torch.float32, device=x.device) if save_stats else None
    check(lib().fs2k_layernorm_fwd(_p(x), _p(_f32(gamma)), _p(_f32(beta)), float(eps), M, D, float(dropout_p), int(seed), _p(y), _p(mean), _p(rstd), _stream()), "fs2k_layernorm_fwd")
    _count()
    return (y, mean, rstd) if save_stats else y


def bn_scale_shift(bn: torch.nn.modules.batchnorm._BatchNorm, z: Optional[torch.Tensor], training: bool, save_stats=False):
    """(scale, shift) of a BatchNorm layer: batch statistics of z[M,C] (and running-stat update) when
    training, folded running statistics otherwise."""
    C = bn.num_features
    dev = bn.weight.device
    scale = torch.empty((C,), dtype=torch.float32, device=dev)
    shift = torch.empty((C,), dtype=torch.float32, device=dev)
    save_mean = torch.empty((C,), dtype=torch.float32, device=dev) if save_stats else None
    save_rstd = torch.empty((C,), dtype=torch.float32, device=dev) if save_stats else None
    sums = None
    M = 0
    if training:
        z = _f32(z, "z")
        M = z.numel() // C
        sums = torch.empty((2 * C,), dtype=torch.float64, device=dev)
        check(lib().fs2k_colstats(_p(z), M, C, _p(sums), _stream()), "fs2k_colstats")
        _count()
    momentum = 0.1 if bn.momentum is None else bn.momentum
    check(lib().fs2k_bn_finalize(_p(sums), M, C, _p(bn.weight), _p(bn.bias), float(bn.eps), float(momentum), int(training),
                                 _p(bn.running_mean), _p(bn.running_var), _p(bn.num_batches_tracked) if training else None,
                                 _p(scale), _p(shift), _p(save_mean), _p(save_rstd), _stream()), "fs2k_bn_finalize")
    _count()
    if save_stats:
        return scale, shift, save_mean, save_rstd
    return scale, shift


def affine_act(z, scale, shift, act=None, residual=None, dropout_p: float = 0.0, seed: int = 0, out_bf16: bool = False):
    z = _f32(z, "z")
    C = z.shape[-1]
    M = z.numel() // C
    if out_bf16:  # the value only feeds a bf16-mode contraction: rounded here, the fp32 copy is never written
        y = torch.empty(z.shape, dtype=torch.bfloat16, device=z.device)
        check(lib().fs2k_affine_act_bf16(_p(z), _p(scale), _p(shift), _ACTS[act], _p(residual), M, C, float(dropout_p), int(seed), _p(y), _stream()),
              "fs2k_affine_act_bf16")
        _count()
        return y
    y = torch.empty_like(z)
    check(lib().fs2k_affine_act(_p(z), _p(scale), _p(shift), _ACTS[act], _p(residual), M, C, float(dropout_p), int(seed), _p(y), _stream()), "fs2k_affine_act")
    _count()
    return y


# ---------------------------------------------------------------------------------------------
# dense contractions
# ---------------------------------------------------------------------------------------------
_repack_cache: dict = {}


def conv_weight_taps(weight: torch.Tensor) -> torch.Tensor:
    """Conv1d weight [N,K,taps] → [taps,N,K] (cached per weight version; 1-tap weights are a view)."""
    N, K, taps = weight.shape
    if taps == 1:
        return weight.detach().reshape(1, N, K)
    # keyed on the tensor OBJECT (weak reference): id()/data_ptr()/_version alone can all be reused by a new
    # parameter allocated at a freed one's address, which would hand back another model's weights
    key = (weight.data_ptr(), weight._version, tuple(weight.shape))
    # while a CUDA graph is being captured the re-pack must be a node of the graph (the weights change between replays)
    # and its output lives in the graph's pool: neither read nor fill the cache then
    capturing = weight.is_cuda and torch.cuda.is_current_stream_capturing()
    hit = None if capturing else _repack_cache.get(id(weight))
    if hit is not None and hit[0] == key and hit[2]() is weight:
        return hit[1]
    w = _f32(weight.detach(), "weight")
    out = torch.empty((taps, N, K), dtype=torch.float32, device=w.device)
    check(lib().fs2k_repack_conv_weight(_p(w), N, K, taps, _p(out), _stream()), "fs2k_repack_conv_weight")
    _count()
    if capturing:
        return out
    if len(_repack_cache) > 512:  # drop entries whose parameter is gone
        for k in [k for k, v in _repack_cache.items() if v[2]() is None]:
            del _repack_cache[k]
    _repack_cache[id(weight)] = (key, out, weakref.ref(weight))
    return out


# Arithmetic of the dense contractions:
#   "fp32"   exact FFMA kernel (gemm_simt.cu)
#   "tf32x3" tcgen05 tensor cores, 3×TF32 split accumulation — fp32-level accuracy (default)
#   "tf32"   tcgen05 tensor cores, single TF32 pass
#   "bf16"   tcgen05 kind::f16: operands rounded to bf16, fp32 accumulation / epilogue / residual stream / norms /
#            softmax / master weights (BASELINE configs[2]).  Applies to the Conformer stacks, mel_linear and PostNet;
#            what feeds a discrete decision (aligner → MAS, variance predictors → bucketize / durations) runs under
#            `full_precision()` and stays 3×TF32, so alignments, durations and teacher-forced bucket ids do not move.
#   "bf16x3" fp32-level accuracy on the bf16 tensor-core path: activations and weights split as hi + lo (two bf16 each),
#            hi·hi + hi·lo + lo·hi accumulated in fp32 (dropped term 2^-16 per product) — 3 kind::f16 MMAs cost half of
#            3 kind::tf32 ones and the operand tiles are half the size.  Weight gradients and attention as in "tf32x3".
PRECISION = "tf32x3"


# Arithmetic of the gradient contractions (dgrad / wgrad); None = same as PRECISION.
BACKWARD_PRECISION = None


# Synthesis only: arithmetic of the contractions AFTER the last discrete decision (decoder, mel_linear, PostNet); None =
# same as PRECISION.  "tf32" there is the reduced-precision configuration north_star allows (mel L1 ≤ 1e-2): bucket ids,
# durations and the length regulator are still decided in fp32-level arithmetic, so nothing flips — only the mel carries
# the ≈1e-3 relative TF32 rounding.
DECODER_PRECISION = None


def set_precision(mode: str, backward: str | None = None, decoder: str | None = None) -> None:
    global PRECISION, BACKWARD_PRECISION, DECODER_PRECISION
    for m in (mode, backward, decoder):
        if m not in (None, "fp32", "tf32", "tf32x3", "bf16", "bf16x3"):
            raise ValueError(m)
    PRECISION = mode
    BACKWARD_PRECISION = backward
    DECODER_PRECISION = decoder


class decoder_precision:
    """Context for the synthesis decoder / mel_linear / PostNet: DECODER_PRECISION when set and no gradient is needed."""

    def __enter__(self):
        global PRECISION
        self.prev = PRECISION
        if DECODER_PRECISION is not None and not torch.is_grad_enabled():
            PRECISION = DECODER_PRECISION

    def __exit__(self, *a):
        global PRECISION
        PRECISION = self.prev


class precision_scope:
    """Run the enclosed contractions in `mode` (the autograd Functions re-enter their forward's mode in backward)."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        global PRECISION
        self.prev = PRECISION
        if self.mode is not None:
            PRECISION = self.mode

    def __exit__(self, *a):
        global PRECISION
        PRECISION = self.prev


class full_precision(precision_scope):
    """fp32-level arithmetic for whatever feeds a discrete decision: "bf16" is lifted to 3×TF32, other modes stay."""

    def __init__(self):
        super().__init__(None)

    def __enter__(self):
        self.mode = "tf32x3" if PRECISION == "bf16" else None
        super().__enter__()


class backward_precision:
    """Context used by the autograd Functions: run the enclosed contractions in BACKWARD_PRECISION."""

    def __enter__(self):
        global PRECISION
        self.prev = PRECISION
        if BACKWARD_PRECISION is not None:
            PRECISION = BACKWARD_PRECISION

    def __exit__(self, *a):
        global PRECISION
        PRECISION = self.prev


def gemm(a, w, bias=None, *, taps_pad: int = 0, scale=None, shift=None, act=None, alpha: float = 1.0,
         residual=None, row_mask=None, out=None, ln=None, ln2=None, want_c: bool = True, dropout_p: float = 0.0,
         seed: int = 0, w_small=None):
    """a [B,L,K] (or [M,K]) · w [taps,N,K] or [N,K] with the fused epilogue of fs2k_gemm_{tc,f32}.

    ln = (gamma, beta, eps): also return LayerNorm(result) — fused into the tensor-core epilogue when one
    tile spans the row, otherwise a separate fs2k_layernorm_fwd launch; ln2 = (gamma, beta) chains a second
    LayerNorm on the first one's output.  Returns C, or (C, ln_out[, ln2_out]) when ln is given."""
    if PRECISION in ("bf16", "bf16x3") and ln is None:
        r = _gemm_bf16_auto(a, w, bias, taps_pad, scale, shift, act, alpha, residual, row_mask, out, dropout_p, seed)
        if r is not None:
            return r
    if ln is not None or PRECISION != "fp32":
        r = _gemm_tc(a, w, bias, taps_pad, scale, shift, act, alpha, residual, row_mask, out, ln, ln2, want_c, dropout_p, seed, w_small)
        if r is not None:
            return r
    if dropout_p:  # SIMT fallback: dropout(+residual) as its own launch
        assert ln is None and row_mask is None
        c = _gemm_f32(a, w, bias, taps_pad=taps_pad, scale=scale, shift=shift, act=act, alpha=alpha)
        return dropout(c, dropout_p, seed, residual)
    c = _gemm_f32(a, w, bias, taps_pad=taps_pad, scale=scale, shift=shift, act=act, alpha=alpha, residual=residual,
                  row_mask=row_mask, out=out)
    if ln is None:
        return c
    y = layernorm(c, ln[0], ln[1], ln[2])
    if ln2 is None:
        return c, y
    return c, y, layernorm(y, ln2[0], ln2[1], ln[2])


def _bf16_shape_ok(a, w, w_mn: bool) -> bool:
    K = a.shape[-1]
    taps = 1 if w.dim() == 2 else w.shape[0]
    N = w.shape[-1] if w_mn else w.shape[-2]
    return bool(lib().fs2k_gemm_bf16_supported(K, N, K, taps, int(a.dtype == torch.bfloat16), int(w_mn))) and not (a.data_ptr() & 15)


def bf16_dgrad_ok(w_taps) -> bool:
    """Can the bf16-mode data-gradient GEMM read these forward weights [taps,N,K] as its MN-major operand?"""
    taps, N, K = w_taps.shape
    return bool(lib().fs2k_gemm_bf16_supported(N, K, N, taps, 0, 1))


def _gemm_bf16_auto(a, w, bias, taps_pad, scale, shift, act, alpha, residual, row_mask, out, dropout_p, seed):
    """The bf16-mode path of gemm(): fp32 weights `w` → their bf16 shadow / cached cast; None when the shape is not taken."""
    if not _bf16_shape_ok(a, w, False):
        return None
    w16, wlo = bf16_weight(w, want_lo=PRECISION == "bf16x3")
    K = a.shape[-1]
    taps = 1 if w.dim() == 2 else w.shape[0]
    hint = 256 if taps * K >= 1024 and w16.shape[-2] % 256 == 0 else 0  # compute-bound shapes (PostNet): wide tiles
    c, _, _ = gemm_bf16(a, w16, bias, w_lo=wlo, taps_pad=taps_pad, scale=scale, shift=shift, act=act, alpha=alpha, residual=residual,
                        row_mask=row_mask, dropout_p=dropout_p, seed=seed, block_n_hint=hint, out=out)
    return c


def gemm_dgrad(g, w_taps, pad: int, wt=None):
    """dX of y = conv(x, W): g [B,L,N] · W [taps,N,K] → [B,L,K].  bf16 mode: the forward weights themselves are the
    (MN-major) operand; otherwise `wt` / a transposed re-pack feeds the forward kernel."""
    taps = w_taps.shape[0]
    if PRECISION in ("bf16", "bf16x3") and bf16_dgrad_ok(w_taps) and not (g.data_ptr() & 15):
        w16, wlo = bf16_weight(w_taps, want_lo=PRECISION == "bf16x3")
        hint = 256 if taps * w_taps.shape[1] >= 1024 and w_taps.shape[2] % 256 == 0 else 0
        return gemm_bf16(g, w16, None, w_lo=wlo, w_mn=True, taps_pad=taps - 1 - pad, block_n_hint=hint)[0]
    return gemm(g, wt if wt is not None else weight_taps_transposed(w_taps), None, taps_pad=taps - 1 - pad)


def gemm_silu_pair(a, w, bias, taps_pad: int = 0, residual=None, dropout_p: float = 0.0, seed: int = 0):
    """(pre, y): pre = conv(a, W) + b, y = dropout(silu(pre)) + residual — one launch in the bf16 mode (the epilogue
    writes both), else the GEMM followed by the fused SiLU + dropout elementwise kernel."""
    if PRECISION in ("bf16", "bf16x3") and _bf16_shape_ok(a, w, False):
        w16, wlo = bf16_weight(w, want_lo=PRECISION == "bf16x3")
        y, _, pre = gemm_bf16(a, w16, bias, w_lo=wlo, taps_pad=taps_pad, act="silu", residual=residual, want_pre="fp32",
                              dropout_p=dropout_p, seed=seed)
        return pre, y
    pre = gemm(a, w, bias, taps_pad=taps_pad)
    return pre, affine_act(pre, None, None, "silu", residual, dropout_p=dropout_p, seed=seed)


def split_small(x):
    """x − tf32(x): the small operand of the 3×TF32 scheme.  `gemm(..., w_small=split_small(w))` lets the kernel load the
    weights' small parts with TMA instead of splitting the weight tile in shared memory at every k-block: 17.6 → 16.3 µs
    per GEMM in a back-to-back chain, but < 1 % on the full training / synthesis steps (measured), so the model code does
    not carry the per-weight caches it would need; the capability stays for callers with static weights."""
    x = _f32(x, "x")
    out = torch.empty_like(x)
    check(lib().fs2k_split_small(_p(x), x.numel(), _p(out), _stream()), "fs2k_split_small")
    _count()
    return out


def _gemm_tc(a, w, bias, taps_pad, scale, shift, act, alpha, residual, row_mask, out, ln, ln2, want_c, dropout_p=0.0, seed=0, w_small=None):
    if PRECISION == "fp32":
        return None
    a = _f32(a, "a")
    if a.dim() == 2:
        B, L, K = 1, a.shape[0], a.shape[1]
        out_shape = (L,)
    else:
        B, L, K = a.shape
        out_shape = (B, L)
    w = _f32(w, "w")
    if w.dim() == 2:
        w = w.reshape(1, *w.shape)
    taps, N, Kw = w.shape
    assert Kw == K, (Kw, K)
    if not lib().fs2k_gemm_tc_supported(K, N, K, taps) or (a.data_ptr() & 15) or (w.data_ptr() & 15):
        return None
    fuse_ln = ln is not None and N <= 256
    dev = a.device
    c = out if out is not None else (torch.empty((*out_shape, N), dtype=torch.float32, device=dev) if (want_c or not fuse_ln) else None)
    if residual is not None:
        residual = _f32(residual, "residual")
    if row_mask is not None:
        row_mask = row_mask.contiguous()
    ln_out = torch.empty((*out_shape, N), dtype=torch.float32, device=dev) if fuse_ln else None
    ln2_out = torch.empty((*out_shape, N), dtype=torch.float32, device=dev) if (fuse_ln and ln2 is not None) else None
    check(lib().fs2k_gemm_tc(_p(a), K, B, L, K, _p(w), N, taps, taps_pad, _p(bias), _p(scale), _p(shift), _ACTS[act],
                             float(alpha), _p(residual), N, _p(row_mask), _p(c), N,
                             _p(ln[0]) if fuse_ln else None, _p(ln[1]) if fuse_ln else None, float(ln[2]) if fuse_ln else 0.0,
                             _p(ln_out), _p(ln2[0]) if ln2_out is not None else None, _p(ln2[1]) if ln2_out is not None else None,
                             _p(ln2_out), float(dropout_p), int(seed), 1 if PRECISION == "tf32" else 3,
                             _p(w_small) if (w_small is not None and PRECISION == "tf32x3") else None, _stream()), "fs2k_gemm_tc")
    _count()
    if ln is None:
        return c
    if not fuse_ln:
        y = layernorm(c, ln[0], ln[1], ln[2])
        return (c, y) if ln2 is None else (c, y, layernorm(y, ln2[0], ln2[1], ln[2]))
    return (c, ln_out) if ln2 is None else (c, ln_out, ln2_out)


def _gemm_f32(a, w, bias=None, *, taps_pad: int = 0, scale=None, shift=None, act=None, alpha: float = 1.0,
              residual=None, row_mask=None, out=None):
    """a [B,L,K] (or [M,K]) · w [taps,N,K] or [N,K] with the fused epilogue of fs2k_gemm_f32."""
    a = _f32(a, "a")
    if a.dim() == 2:
        B, L, K = 1, a.shape[0], a.shape[1]
        out_shape = (L,)
    else:
        B, L, K = a.shape
        out_shape = (B, L)
    w = _f32(w, "w")
    if w.dim() == 2:
        w = w.reshape(1, *w.shape)
    taps, N, Kw = w.shape
    assert Kw == K, (Kw, K)
    c = out if out is not None else torch.empty((*out_shape, N), dtype=torch.float32, device=a.device)
    if residual is not None:
        residual = _f32(residual, "residual")
    if row_mask is not None:
        row_mask = row_mask.contiguous()
        assert row_mask.dtype in (torch.bool, torch.uint8)
    check(lib().fs2k_gemm_f32(_p(a), K, B, L, K, _p(w), N, taps, taps_pad, _p(bias), _p(scale), _p(shift), _ACTS[act],
                              float(alpha), _p(residual), N, _p(row_mask), _p(c), N, _stream()), "fs2k_gemm_f32")
    _count()
    return c


# ---------------------------------------------------------------------------------------------
# bf16 arithmetic mode (gemm_bf16.cu / gemm_wgrad_bf16.cu)
# ---------------------------------------------------------------------------------------------
_bf16_shadows: list = []   # weak references to FusedAdamW instances that keep a bf16 copy of their flat parameter buffer
_bf16_cache: dict = {}


def register_bf16_shadow(optimizer) -> None:
    _bf16_shadows[:] = [r for r in _bf16_shadows if r() is not None and r() is not optimizer]
    _bf16_shadows.append(weakref.ref(optimizer))


def refresh_bf16_shadows(parameters=None) -> None:
    """Once per model forward in the bf16 mode: re-sync the optimizers' bf16 weight shadows if anything but the update
    kernel wrote the parameters since.  `parameters` (an iterable of the model's parameters) creates the shadow of the
    optimizer that owns them on first use."""
    if parameters is not None:
        first = next(iter(parameters), None)
        owner = getattr(first, "_fs2k_flat_owner", None) if first is not None else None
        owner = owner() if owner is not None else None
        if owner is not None:
            owner.bf16_shadow()
            return
    for r in _bf16_shadows:
        opt = r()
        if opt is not None:
            opt.bf16_shadow()


def bf16_weight(w: torch.Tensor, want_lo: bool = False):
    """(hi, lo) bf16 operands of an fp32 weight tensor: a view of the owning optimizer's shadow buffer when there is one
    (zero launches), else a cast cached per weight version (static weights: synthesis / validation), else a cast."""
    ptr = w.data_ptr()
    if not want_lo:
        for r in _bf16_shadows:
            opt = r()
            if opt is None or opt.flat_p16 is None:
                continue
            base = opt.flat_p.data_ptr()
            if base <= ptr < base + 4 * opt.flat_p.numel() and w.is_contiguous():
                off = (ptr - base) // 4
                return opt.flat_p16[off: off + w.numel()].view(w.shape), None
    capturing = torch.cuda.is_current_stream_capturing()
    key = (ptr, w._version, tuple(w.shape), want_lo)
    hit = None if capturing else _bf16_cache.get(ptr)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    hi, lo = cast_bf16(w, want_lo)
    if not capturing:
        if len(_bf16_cache) > 256:
            _bf16_cache.clear()
        _bf16_cache[ptr] = (key, hi, lo, w)  # holding w keeps its address from being reused while the entry lives
    return hi, lo


def cast_bf16(x, want_lo: bool = False):
    """fp32 → (hi = bf16(x), lo = bf16(x − hi) or None)."""
    x = _f32(x, "x")
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_lo else None
    check(lib().fs2k_cast_bf16(_p(x), x.numel(), _p(hi), _p(lo), _stream()), "fs2k_cast_bf16")
    _count()
    return hi, lo


def gemm_bf16(a, w_hi, bias=None, *, w_lo=None, w_mn: bool = False, taps_pad: int = 0, scale=None, shift=None, act=None,
              alpha: float = 1.0, residual=None, row_mask=None, want_c: bool = True, want_c16: bool = False,
              want_pre: str | None = None, dropout_p: float = 0.0, seed: int = 0, block_n_hint: int = 0, out=None):
    """a [B,L,K] / [M,K] (fp32, rounded in the kernel, or bf16) · bf16 weights, fp32 accumulate + fp32 epilogue.

    w_mn=False: w_hi [taps,N,K] (or [N,K]).  w_mn=True: w_hi [taps,K,N] read MN-major, taps reversed (data gradient).
    w_lo: bf16(W − w_hi) → hi·hi + hi·lo + lo·hi.  Returns (C fp32 | None, C16 bf16 | None, pre | None)."""
    if not a.is_cuda:
        raise ValueError("a must be a CUDA tensor (no CPU path)")
    a_is_bf16 = a.dtype == torch.bfloat16
    if not a_is_bf16:
        a = _f32(a, "a")
    elif not a.is_contiguous():
        a = a.contiguous()
    if a.dim() == 2:
        B, L, K = 1, a.shape[0], a.shape[1]
        out_shape = (L,)
    else:
        B, L, K = a.shape
        out_shape = (B, L)
    assert w_hi.dtype == torch.bfloat16 and w_hi.is_contiguous()
    if w_hi.dim() == 2:
        w_hi = w_hi.reshape(1, *w_hi.shape)
        if w_lo is not None:
            w_lo = w_lo.reshape(1, *w_lo.shape)
    taps = w_hi.shape[0]
    if w_mn:
        assert w_hi.shape[1] == K, (w_hi.shape, K)
        N = w_hi.shape[2]
    else:
        assert w_hi.shape[2] == K, (w_hi.shape, K)
        N = w_hi.shape[1]
    if not lib().fs2k_gemm_bf16_supported(K, N, K, taps, int(a_is_bf16), int(w_mn)):
        raise ValueError(f"fs2k_gemm_bf16 does not take K={K} N={N} taps={taps} bf16_a={a_is_bf16} w_mn={w_mn}")
    dev = a.device
    c = out if out is not None else (torch.empty((*out_shape, N), dtype=torch.float32, device=dev) if want_c else None)
    c16 = torch.empty((*out_shape, N), dtype=torch.bfloat16, device=dev) if want_c16 else None
    pre = None
    if want_pre is not None:
        pre = torch.empty((*out_shape, N), dtype=torch.float32 if want_pre == "fp32" else torch.bfloat16, device=dev)
    if residual is not None:
        residual = _f32(residual, "residual")
    if row_mask is not None:
        row_mask = row_mask.contiguous()
    check(lib().fs2k_gemm_bf16(_p(a), int(a_is_bf16), K, B, L, K, _p(w_hi), _p(w_lo), int(w_mn), N, taps, taps_pad, _p(bias),
                               _p(scale), _p(shift), _ACTS[act], float(alpha), _p(residual), N, _p(row_mask), _p(c), N, _p(c16), N,
                               _p(pre) if want_pre == "fp32" else None, _p(pre) if want_pre == "bf16" else None, N,
                               float(dropout_p), int(seed), int(block_n_hint), _stream()), "fs2k_gemm_bf16")
    _count()
    return c, c16, pre


def gemm_bf16_dact(g, w16, pre16, act: str, *, w_mn: bool = True, alpha: float = 1.0, dropout_p: float = 0.0, seed: int = 0,
                   want_c: bool = False, want_c16: bool = True):
    """(g · W) ∘ act'(pre16) ∘ dropout mask · alpha — the data-gradient GEMM of a Linear+activation(+dropout) layer fused with the
    activation's derivative.  g [M,K] fp32 (K ≤ 256), w16 [K,N] bf16 (w_mn) or [N,K], pre16 [M,N] bf16 = the forward's saved
    pre-activation.  Returns (C fp32 | None, C16 bf16 | None)."""
    g = _f32(g, "g")
    K = g.shape[-1]
    M = g.numel() // K
    w2 = w16.reshape(w16.shape[-2], w16.shape[-1])
    N = w2.shape[1] if w_mn else w2.shape[0]
    assert pre16.dtype == torch.bfloat16 and pre16.is_contiguous() and pre16.numel() == M * N
    c = torch.empty((*g.shape[:-1], N), dtype=torch.float32, device=g.device) if want_c else None
    c16 = torch.empty((*g.shape[:-1], N), dtype=torch.bfloat16, device=g.device) if want_c16 else None
    check(lib().fs2k_gemm_bf16_dact(_p(g), K, M, K, _p(w2), int(w_mn), N, _p(pre16), _ACTS[act], float(alpha), float(dropout_p), int(seed),
                                    _p(c), _p(c16), _stream()), "fs2k_gemm_bf16_dact")
    _count()
    return c, c16


def bf16_dact_ok(K: int, N: int) -> bool:
    return K % 64 == 0 and K <= 256 and N % 128 == 0


def _f32_or_bf16(t, name):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU path)")
    if t.dtype == torch.bfloat16:
        return t if t.is_contiguous() else t.contiguous()
    return _f32(t, name)


def gemm_wgrad_bf16(g, x, taps: int, pad: int, conv_layout: bool, accumulate_into=None):
    """bf16-mode weight gradient: g [B,L,N], x [B,L,K] (each fp32 — rounded in the kernel — or bf16 — read by TMA) →
    [N,K,taps] (conv_layout) or [N,K]; None when the shape is not taken."""
    g, x = _f32_or_bf16(g, "g"), _f32_or_bf16(x, "x")
    if g.dim() == 2:
        B, L = 1, g.shape[0]
    else:
        B, L = g.shape[0], g.shape[1]
    N, K = g.shape[-1], x.shape[-1]
    g16, x16 = g.dtype == torch.bfloat16, x.dtype == torch.bfloat16
    if B * L == 0 or not lib().fs2k_gemm_wgrad_bf16_supported(N, K, N, K) or (g16 and N % 8) or (x16 and K % 8):
        return None
    ws_bytes = lib().fs2k_gemm_wgrad_bf16_workspace_bytes(B, L, N, K, taps)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g.device)
    acc = accumulate_into is not None and accumulate_into.is_contiguous() and accumulate_into.dtype == torch.float32
    out = accumulate_into if acc else torch.empty((N, K, taps) if conv_layout else (N, K), dtype=torch.float32, device=g.device)
    assert out.numel() == N * K * taps
    check(lib().fs2k_gemm_wgrad_bf16_ex(_p(g), int(g16), N, _p(x), int(x16), K, B, L, N, K, taps, pad, _p(ws), ws_bytes, _p(out), int(acc),
                                        _stream()), "fs2k_gemm_wgrad_bf16_ex")
    _count(2)
    if accumulate_into is not None and not acc:
        accumulate_into.add_(out)
    return out


def rowdot(x, w, b=None, mask=None):
    x = _f32(x, "x")
    D = x.shape[-1]
    M = x.numel() // D
    y = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
    if mask is not None:
        mask = mask.contiguous()
    check(lib().fs2k_rowdot(_p(x), _p(_f32(w)), _p(b), _p(mask), M, D, _p(y), _stream()), "fs2k_rowdot")
    _count()
    return y


def attention_order(lens):
    """Utterance indices, longest first (int32 [B]) — dispatch order of the attention CTAs (work ∝ length)."""
    lens = _i32(lens, "lens")
    B = lens.numel()
    if B < 2 or B > 4096:
        return None
    order = torch.empty((B,), dtype=torch.int32, device=lens.device)
    check(lib().fs2k_attention_order(_p(lens), B, _p(order), _stream()), "fs2k_attention_order")
    _count()
    return order


def attention(qkv, lens, heads: int, want_lse: bool = False, dropout_p: float = 0.0, seed: int = 0, order=None):
    qkv, lens = _f32(qkv, "qkv"), _i32(lens, "lens")
    B, L, D3 = qkv.shape
    D = D3 // 3
    out = torch.empty((B, L, D), dtype=torch.float32, device=qkv.device)
    lse = torch.empty((B, heads, L), dtype=torch.float32, device=qkv.device) if want_lse else None
    check(lib().fs2k_attention_f32(_p(qkv), _p(lens), B, L, heads, D // heads, float(dropout_p), int(seed), _p(out), _p(lse), _p(order), _stream()), "fs2k_attention_f32")
    _count()
    return (out, lse) if want_lse else out


def attention_bf16(qkv16, lens, heads: int, want_lse: bool = False, dropout_p: float = 0.0, seed: int = 0, order=None):
    """Tensor-core attention of the bf16 mode: qkv16 [B,L,3D] bf16 → out [B,L,D] fp32 (+ lse [B,H,L])."""
    assert qkv16.is_cuda and qkv16.dtype == torch.bfloat16 and qkv16.is_contiguous()
    lens = _i32(lens, "lens")
    B, L, D3 = qkv16.shape
    D = D3 // 3
    out = torch.empty((B, L, D), dtype=torch.float32, device=qkv16.device)
    lse = torch.empty((B, heads, L), dtype=torch.float32, device=qkv16.device) if want_lse else None
    check(lib().fs2k_attention_bf16(_p(qkv16), _p(lens), B, L, heads, D // heads, float(dropout_p), int(seed), _p(out), _p(lse), _p(order), _stream()),
          "fs2k_attention_bf16")
    _count()
    return (out, lse) if want_lse else out


def attention_bwd_bf16(qkv16, out, lse, dout, lens, heads: int, dropout_p: float = 0.0, seed: int = 0, order=None):
    assert qkv16.dtype == torch.bfloat16 and qkv16.is_contiguous()
    out, dout = _f32(out, "out"), _f32(dout, "dout")
    B, L, D3 = qkv16.shape
    D = D3 // 3
    dout16, _ = cast_bf16(dout)
    delta = torch.empty((B, heads, L), dtype=torch.float32, device=qkv16.device)
    dqkv = torch.empty((B, L, D3), dtype=torch.float32, device=qkv16.device)
    check(lib().fs2k_attention_bwd_bf16(_p(qkv16), _p(out), _p(lse), _p(dout), _p(dout16), _p(_i32(lens)), B, L, heads, D // heads,
                                        float(dropout_p), int(seed), _p(delta), _p(dqkv), _p(order), _stream()), "fs2k_attention_bwd_bf16")
    _count(3)
    return dqkv


def dwconv(x, weight, bias, *, channels: int, glu: bool = False, scale=None, shift=None):
    """x [B,L,ldx] → [B,L,channels]; weight [C,1,K]."""
    x = _f32(x, "x")
    B, L, ldx = x.shape
    K = weight.shape[-1]
    y = torch.empty((B, L, channels), dtype=torch.float32, device=x.device)
    check(lib().fs2k_dwconv_fwd(_p(x), ldx, B, L, channels, _p(_f32(weight)), K, _p(bias), int(glu), _p(scale), _p(shift), _p(y), _stream()), "fs2k_dwconv_fwd")
    _count()
    return y


def aligner_scores(q, k, prior, key_lens):
    q, k = _f32(q, "q"), _f32(k, "k")
    B, F, C = q.shape
    T = k.shape[1]
    if prior is not None:
        prior = _f32(prior, "prior")
        assert prior.shape == (B, F, T), (prior.shape, (B, F, T))
    logprob = torch.empty((B, 1, F, T), dtype=torch.float32, device=q.device)
    soft = torch.empty((B, 1, F, T), dtype=torch.float32, device=q.device)
    kl = _i32(key_lens, "key_lens") if key_lens is not None else None
    check(lib().fs2k_aligner_fwd(_p(q), _p(k), _p(prior), _p(kl), B, F, T, C, _p(logprob), _p(soft), _stream()), "fs2k_aligner_fwd")
    _count(2)
    return soft, logprob


# ---------------------------------------------------------------------------------------------
# losses and small elementwise helpers
# ---------------------------------------------------------------------------------------------
_KIND = {"mse": 0, "mae": 1}


def masked_loss_fwd(pred, target, row_mask, kind: str, weight: float, log1p_int_target: bool = False):
    pred = _f32(pred, "pred")
    target = _i32(target, "target") if log1p_int_target else _f32(target, "target")
    row_mask = row_mask.contiguous()
    M = row_mask.numel()
    C = pred.numel() // max(M, 1)
    assert pred.numel() == target.numel() == M * C, (pred.shape, target.shape, row_mask.shape)
    scratch = torch.empty((1,), dtype=torch.float64, device=pred.device)
    loss = torch.empty((), dtype=torch.float32, device=pred.device)
    check(lib().fs2k_masked_loss_fwd(_p(pred), _p(target), int(log1p_int_target), _p(row_mask), M, C, _KIND[kind], float(weight), _p(scratch), _p(loss), _stream()), "fs2k_masked_loss_fwd")
    _count(2)
    return loss


def masked_loss_bwd(pred, target, row_mask, kind: str, weight: float, gout, log1p_int_target: bool = False):
    pred = _f32(pred, "pred")
    target = _i32(target, "target") if log1p_int_target else _f32(target, "target")
    row_mask = row_mask.contiguous()
    M = row_mask.numel()
    C = pred.numel() // max(M, 1)
    dpred = torch.empty_like(pred)
    check(lib().fs2k_masked_loss_bwd(_p(pred), _p(target), int(log1p_int_target), _p(row_mask), M, C, _KIND[kind], float(weight), _p(_f32(gout)), _p(dpred), _stream()), "fs2k_masked_loss_bwd")
    _count()
    return dpred


def bin_loss_fwd(hard, soft, eps: float):
    hard, soft = _f32(hard, "hard"), _f32(soft, "soft")
    sums = torch.empty((2,), dtype=torch.float64, device=soft.device)
    loss = torch.empty((), dtype=torch.float32, device=soft.device)
    check(lib().fs2k_bin_loss_fwd(_p(hard), _p(soft), soft.numel(), float(eps), _p(sums), _p(loss), _stream()), "fs2k_bin_loss_fwd")
    _count(2)
    return loss, sums


def bin_loss_bwd(hard, soft, eps: float, sums, gout):
    dsoft = torch.empty_like(soft)
    check(lib().fs2k_bin_loss_bwd(_p(hard), _p(soft), soft.numel(), float(eps), _p(sums), _p(_f32(gout)), _p(dsoft), _stream()), "fs2k_bin_loss_bwd")
    _count()
    return dsoft


def ctc_forward_sum_fwd(attn_logprob, key_lens, query_lens, blank_logprob: float = -1.0):
    """Forward-sum loss of attn_logprob [B,1,F,T] (attention_loss.py:22-62) → (loss 0-d, saved tensors for the backward)."""
    x = _f32(attn_logprob, "attn_logprob")
    B, F, T = x.shape[0], x.shape[-2], x.shape[-1]
    kl, ql = _i32(key_lens, "key_lens"), _i32(query_lens, "query_lens")
    dev = x.device
    lse = torch.empty((B, F), dtype=torch.float64, device=dev)
    log_alpha = torch.empty((B, F, 2 * T + 1), dtype=torch.float64, device=dev)
    nll = torch.empty((B,), dtype=torch.float64, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    check(lib().fs2k_ctc_forward_sum_fwd(_p(x), _p(kl), _p(ql), B, F, T, float(blank_logprob), _p(lse), _p(log_alpha), _p(nll), _p(loss), _stream()),
          "fs2k_ctc_forward_sum_fwd")
    _count(2)
    return loss, (x, lse, log_alpha, nll, kl, ql)


def ctc_forward_sum_bwd(saved, gout, blank_logprob: float = -1.0):
    x, lse, log_alpha, nll, kl, ql = saved
    B, F, T = x.shape[0], x.shape[-2], x.shape[-1]
    dx = torch.empty_like(x)
    check(lib().fs2k_ctc_forward_sum_bwd(_p(x), _p(lse), _p(log_alpha), _p(nll), _p(kl), _p(ql), _p(_f32(gout)), B, F, T, float(blank_logprob), _p(dx), _stream()),
          "fs2k_ctc_forward_sum_bwd")
    _count()
    return dx


def conv2d_s2_bn_relu(x, w_packed, scale, shift, cw_layout: bool = False):
    """One GST ReferenceEncoder block on channels-last x [B,H,W,Ci]; w_packed [3,3,Ci,Co] → [B,Ho,Wo,Co] ([B,Ho,Co,Wo])."""
    x = _f32(x, "x")
    B, H, W, Ci = x.shape
    Co = w_packed.shape[-1]
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((B, Ho, Co, Wo) if cw_layout else (B, Ho, Wo, Co), dtype=torch.float32, device=x.device)
    check(lib().fs2k_conv2d_s2_bn_relu(_p(x), _p(_f32(w_packed, "w")), _p(scale), _p(shift), B, H, W, Ci, Co, int(cw_layout),
                                       _p(y), _stream()), "fs2k_conv2d_s2_bn_relu")
    _count()
    return y


def conv2d_s2_raw(x, w_packed):
    """Pre-BatchNorm output of a ReferenceEncoder block (training path): [B,H,W,Ci] → [B,Ho,Wo,Co]."""
    x = _f32(x, "x")
    B, H, W, Ci = x.shape
    Co = w_packed.shape[-1]
    y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, Co), dtype=torch.float32, device=x.device)
    check(lib().fs2k_conv2d_s2_bn_relu(_p(x), _p(_f32(w_packed, "w")), None, None, B, H, W, Ci, Co, 0, _p(y), _stream()),
          "fs2k_conv2d_s2_bn_relu")
    _count()
    return y


def conv2d_s2_dgrad(gz, w_packed, x_shape):
    gz = _f32(gz, "gz")
    B, H, W, Ci = x_shape
    Co = w_packed.shape[-1]
    dx = torch.empty(tuple(x_shape), dtype=torch.float32, device=gz.device)
    check(lib().fs2k_conv2d_s2_dgrad(_p(gz), _p(_f32(w_packed, "w")), B, H, W, Ci, Co, _p(dx), _stream()), "fs2k_conv2d_s2_dgrad")
    _count()
    return dx


def conv2d_s2_wgrad(x, gz):
    """[3,3,Ci,Co] weight gradient of the stride-2 3×3 convolution."""
    x, gz = _f32(x, "x"), _f32(gz, "gz")
    B, H, W, Ci = x.shape
    Co = gz.shape[-1]
    dw = torch.empty((3, 3, Ci, Co), dtype=torch.float32, device=x.device)
    check(lib().fs2k_conv2d_s2_wgrad(_p(x), _p(gz), B, H, W, Ci, Co, _p(dw), _stream()), "fs2k_conv2d_s2_wgrad")
    _count()
    return dw


def gru_gate(xp, t: int, T: int, hp, h):
    """One GRU cell update from xp [B·T, 3U] (step t of T) and hp = W_hh·h + b_hh."""
    B, U = h.shape
    h_new = torch.empty_like(h)
    check(lib().fs2k_gru_gate(xp.data_ptr() + 4 * t * 3 * U, T * 3 * U, _p(hp), _p(h), B, U, _p(h_new), _stream()), "fs2k_gru_gate")
    _count()
    return h_new


def gru_gate_bwd(xp, dxp, t: int, T: int, hp, h, dh):
    """Cell backward of step t: fills dxp[:, t] and returns (d_hp [B,3U], dh·z [B,U])."""
    B, U = h.shape
    dhp = torch.empty_like(hp)
    dh_prev = torch.empty_like(h)
    off = 4 * t * 3 * U
    check(lib().fs2k_gru_gate_bwd(xp.data_ptr() + off, T * 3 * U, _p(hp), _p(h), _p(_f32(dh, "dh")), B, U, dxp.data_ptr() + off,
                                  T * 3 * U, _p(dhp), _p(dh_prev), _stream()), "fs2k_gru_gate_bwd")
    _count()
    return dhp, dh_prev


def gst_token_attention_bwd(q, k, v, dout, heads: int):
    q, k, v, dout = _f32(q, "q"), _f32(k, "k"), _f32(v, "v"), _f32(dout, "dout")
    B, D = q.shape
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    check(lib().fs2k_gst_token_attention_bwd(_p(q), _p(k), _p(v), _p(dout), B, k.shape[0], heads, D // heads, _p(dq), _p(dk), _p(dv),
                                             _stream()), "fs2k_gst_token_attention_bwd")
    _count()
    return dq, dk, dv


def gru_last_hidden(x, w_ih, w_hh, b_ih, b_hh):
    """Last hidden state [B,U] of a one-layer batch_first torch.nn.GRU over x [B,T,I], h0 = 0."""
    x = _f32(x, "x")
    B, T, I = x.shape
    U = w_hh.shape[1]
    xp = gemm(x.reshape(B * T, I), w_ih, b_ih)  # [B·T, 3U]: input projection of every step in one GEMM
    h = torch.zeros((B, U), dtype=torch.float32, device=x.device)
    for t in range(T):
        hp = gemm(h, w_hh, b_hh)
        h_new = torch.empty_like(h)
        check(lib().fs2k_gru_gate(xp.data_ptr() + 4 * t * 3 * U, T * 3 * U, _p(hp), _p(h), B, U, _p(h_new), _stream()), "fs2k_gru_gate")
        _count()
        h = h_new
    return h


def gst_token_attention(q, k, v, heads: int):
    q, k, v = _f32(q, "q"), _f32(k, "k"), _f32(v, "v")
    B, D = q.shape
    out = torch.empty_like(q)
    check(lib().fs2k_gst_token_attention(_p(q), _p(k), _p(v), B, k.shape[0], heads, D // heads, _p(out), _stream()), "fs2k_gst_token_attention")
    _count()
    return out


def axpby(a, alpha: float, b=None, beta: float = 0.0):
    a = _f32(a, "a")
    if b is not None:
        b = _f32(b, "b")
        assert b.shape == a.shape, (a.shape, b.shape)
    out = torch.empty_like(a)
    check(lib().fs2k_axpby(_p(a), float(alpha), _p(b), float(beta), a.numel(), _p(out), _stream()), "fs2k_axpby")
    _count()
    return out


def gather_rows(table, ids):
    table = _f32(table, "table")
    ids = ids.contiguous()
    assert ids.dtype == torch.int64
    out = torch.empty((*ids.shape, table.shape[1]), dtype=torch.float32, device=table.device)
    check(lib().fs2k_gather_rows(_p(table), _p(ids), ids.numel(), table.shape[1], _p(out), _stream()), "fs2k_gather_rows")
    _count()
    return out


def dropout(x, p: float, seed: int, residual=None):
    x = _f32(x, "x")
    y = torch.empty_like(x)
    check(lib().fs2k_dropout(_p(x), _p(residual), float(p), int(seed), x.numel(), _p(y), _stream()), "fs2k_dropout")
    _count()
    return y


def tanh(x):
    x = _f32(x, "x")
    y = torch.empty_like(x)
    check(lib().fs2k_tanh(_p(x), x.numel(), _p(y), _stream()), "fs2k_tanh")
    _count()
    return y


# ---------------------------------------------------------------------------------------------
# backward kernels
# ---------------------------------------------------------------------------------------------
_ACT_BWD_MODE = {None: 0, "none": 0, "relu": 1, "silu": 2, "tanh": 3}


def act_bwd(g, aux, act, alpha: float = 1.0, row_mask=None, dropout_p: float = 0.0, seed: int = 0):
    g = _f32(g, "g")
    C = g.shape[-1]
    M = g.numel() // C
    gz = torch.empty_like(g)
    if row_mask is not None:
        row_mask = row_mask.contiguous()
    check(lib().fs2k_act_bwd(_p(g), _p(aux), _ACT_BWD_MODE[act], float(alpha), _p(row_mask), M, C, float(dropout_p), int(seed), _p(gz), _stream()), "fs2k_act_bwd")
    _count()
    return gz


def colsum(z, out=None, accumulate: bool = False):
    """out[c] (+)= Σ_m z[m,c]  (bias gradients); `out` + `accumulate` add straight into an existing gradient."""
    z = _f32_or_bf16(z, "z")
    C = z.shape[-1]
    if out is None:
        out, accumulate = torch.empty((C,), dtype=torch.float32, device=z.device), False
    assert out.is_contiguous() and out.numel() == C and out.dtype == torch.float32
    if z.dtype == torch.bfloat16:
        check(lib().fs2k_colsum_bf16(_p(z), z.numel() // C, C, _p(out), int(accumulate), _stream()), "fs2k_colsum_bf16")
        _count()
        return out
    check(lib().fs2k_colsum(_p(z), z.numel() // C, C, _p(out), int(accumulate), _stream()), "fs2k_colsum")
    _count()
    return out


def weight_taps_transposed(w_taps):
    """[taps,N,K] → [taps,K,N] with reversed taps (weights of the transposed convolution)."""
    w_taps = _f32(w_taps, "w_taps")
    taps, N, K = w_taps.shape
    out = torch.empty((taps, K, N), dtype=torch.float32, device=w_taps.device)
    check(lib().fs2k_repack_weight_t(_p(w_taps), N, K, taps, _p(out), _stream()), "fs2k_repack_weight_t")
    _count()
    return out


def gemm_wgrad(g, x, taps: int, pad: int, conv_layout: bool, accumulate_into=None):
    """dW for y = conv(x, W): g [B,L,N], x [B,L,K] → [N,K,taps] (conv_layout) or [N,K].  `accumulate_into` (a
    contiguous fp32 tensor of that shape, e.g. the parameter's .grad) receives `+= dW` instead of a new tensor."""
    g, x = _f32_or_bf16(g, "g"), _f32_or_bf16(x, "x")
    if g.dim() == 2:
        B, L = 1, g.shape[0]
    else:
        B, L = g.shape[0], g.shape[1]
    N, K = g.shape[-1], x.shape[-1]
    if PRECISION == "bf16":
        out = gemm_wgrad_bf16(g, x, taps, pad, conv_layout, accumulate_into)
        if out is not None:
            return out
        if g.dtype == torch.bfloat16 or x.dtype == torch.bfloat16:
            g, x = g.float(), x.float()
    if PRECISION != "fp32" and B * L > 0 and lib().fs2k_gemm_wgrad_tc_supported(N, K, N, K):
        # tensor cores (tcgen05, MN-major operands); the kernel writes the parameter layout directly
        ws_bytes = lib().fs2k_gemm_wgrad_tc_workspace_bytes(B, L, N, K, taps)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g.device)
        acc = accumulate_into is not None and accumulate_into.is_contiguous() and accumulate_into.dtype == torch.float32
        out = accumulate_into if acc else torch.empty((N, K, taps) if conv_layout else (N, K), dtype=torch.float32, device=g.device)
        assert out.numel() == N * K * taps
        check(lib().fs2k_gemm_wgrad_tc(_p(g), N, _p(x), K, B, L, N, K, taps, pad, 1 if PRECISION == "tf32" else 3,
                                       _p(ws), ws_bytes, _p(out), int(acc), _stream()), "fs2k_gemm_wgrad_tc")
        _count(2)
        if accumulate_into is not None and not acc:
            accumulate_into.add_(out)
        return out
    dw = torch.empty((taps, N, K), dtype=torch.float32, device=g.device)
    check(lib().fs2k_gemm_wgrad(_p(g), N, _p(x), K, B, L, N, K, taps, pad, _p(dw), _stream()), "fs2k_gemm_wgrad")
    _count()
    if not conv_layout:
        out = dw[0]
    elif taps == 1:
        out = dw.reshape(N, K, 1)
    else:
        out = torch.empty((N, K, taps), dtype=torch.float32, device=g.device)
        check(lib().fs2k_unpack_conv_weight(_p(dw), N, K, taps, _p(out), _stream()), "fs2k_unpack_conv_weight")
        _count()
    if accumulate_into is not None:
        accumulate_into.add_(out)
    return out


def _grad_targets(into, shapes, device):
    """(buffers, accumulate): `into` = parameter-gradient tensors to add onto (contiguous fp32), or None → fresh ones."""
    if into is not None and all(t is not None and t.is_contiguous() and t.dtype == torch.float32 for t in into):
        return list(into), 1
    return [torch.empty(s, dtype=torch.float32, device=device) for s in shapes], 0


def layernorm_bwd(g, x, mean, rstd, gamma, dropout_p: float = 0.0, seed: int = 0, accumulate_into=None, add=None):
    """`add`: gradient of the residual branch (same shape as x) summed into dx by the same launch."""
    g, x = _f32(g, "g"), _f32(x, "x")
    D = x.shape[-1]
    M = x.numel() // D
    dx = torch.empty_like(x)
    (dgamma, dbeta), acc = _grad_targets(accumulate_into, [(D,), (D,)], x.device)
    if add is not None:
        assert not dropout_p
        check(lib().fs2k_layernorm_bwd_add(_p(g), _p(x), _p(mean), _p(rstd), _p(_f32(gamma)), M, D, _p(_f32(add, "add")), _p(dx), _p(dgamma), _p(dbeta),
                                           acc, _stream()), "fs2k_layernorm_bwd_add")
        _count()
        return dx, dgamma, dbeta
    check(lib().fs2k_layernorm_bwd(_p(g), _p(x), _p(mean), _p(rstd), _p(_f32(gamma)), M, D, float(dropout_p), int(seed), _p(dx), _p(dgamma), _p(dbeta), acc, _stream()), "fs2k_layernorm_bwd")
    _count()
    return dx, dgamma, dbeta


def bn_act_bwd(g, z, scale, shift, mean, rstd, act, training: bool, dropout_p: float = 0.0, seed: int = 0, accumulate_into=None,
               out_bf16: bool = False):
    g, z = _f32(g, "g"), _f32(z, "z")
    C = z.shape[-1]
    M = z.numel() // C
    sums = torch.empty((2 * C,), dtype=torch.float64, device=z.device)
    (dgamma, dbeta), acc = _grad_targets(accumulate_into, [(C,), (C,)], z.device)
    if out_bf16:
        gz = torch.empty(z.shape, dtype=torch.bfloat16, device=z.device)
        check(lib().fs2k_bn_act_bwd_bf16(_p(g), _p(z), _p(scale), _p(shift), _p(mean), _p(rstd), _ACTS[act], int(training), M, C,
                                         float(dropout_p), int(seed), _p(sums), _p(gz), _p(dgamma), _p(dbeta), acc, _stream()), "fs2k_bn_act_bwd_bf16")
        _count(2)
        return gz, dgamma, dbeta
    gz = torch.empty_like(z)
    check(lib().fs2k_bn_act_bwd(_p(g), _p(z), _p(scale), _p(shift), _p(mean), _p(rstd), _ACTS[act], int(training), M, C,
                                float(dropout_p), int(seed), _p(sums), _p(gz), _p(dgamma), _p(dbeta), acc, _stream()), "fs2k_bn_act_bwd")
    _count(2)
    return gz, dgamma, dbeta


def attention_bwd(qkv, out, lse, dout, lens, heads: int, dropout_p: float = 0.0, seed: int = 0, order=None):
    qkv, out, dout = _f32(qkv, "qkv"), _f32(out, "out"), _f32(dout, "dout")
    B, L, D3 = qkv.shape
    D = D3 // 3
    delta = torch.empty((B, heads, L), dtype=torch.float32, device=qkv.device)
    dqkv = torch.empty_like(qkv)
    check(lib().fs2k_attention_bwd_f32(_p(qkv), _p(out), _p(lse), _p(dout), _p(_i32(lens)), B, L, heads, D // heads, float(dropout_p), int(seed), _p(delta), _p(dqkv), _p(order), _stream()), "fs2k_attention_bwd_f32")
    _count(3)
    return dqkv


def dwconv_bwd(gz, x, weight, glu: bool, want_bias: bool = True, accumulate_into=None):
    gz, x = _f32(gz, "gz"), _f32(x, "x")
    B, L, ldx = x.shape
    C, _, K = weight.shape
    dx = torch.empty_like(x)
    if accumulate_into is not None and want_bias:
        (dw, db), acc = _grad_targets(accumulate_into, [tuple(weight.shape), (C,)], x.device)
    else:
        dw, acc = torch.empty_like(weight), 0
        db = torch.empty((C,), dtype=torch.float32, device=x.device) if want_bias else None
    check(lib().fs2k_dwconv_bwd(_p(gz), _p(x), ldx, B, L, C, _p(_f32(weight)), K, int(glu), _p(dx), _p(dw), _p(db), acc, _stream()), "fs2k_dwconv_bwd")
    _count()
    return dx, dw, db


def rowdot_bwd(g, x, w, mask):
    g, x = _f32(g, "g"), _f32(x, "x")
    D = x.shape[-1]
    M = x.numel() // D
    dx = torch.empty_like(x)
    dw = torch.empty((D,), dtype=torch.float32, device=x.device)
    db = torch.empty((1,), dtype=torch.float32, device=x.device)
    if mask is not None:
        mask = mask.contiguous()
    check(lib().fs2k_rowdot_bwd(_p(g), _p(x), _p(_f32(w)), _p(mask), M, D, _p(dx), _p(dw), _p(db), _stream()), "fs2k_rowdot_bwd")
    _count()
    return dx, dw, db


def lr_bwd(g_out, g_out_pos, cum, T: int, D: int, F_out: int):
    B = cum.shape[0]
    ref = g_out if g_out is not None else g_out_pos
    dx = torch.empty((B, T, D), dtype=torch.float32, device=ref.device)
    g1 = _f32(g_out) if g_out is not None else None
    g2 = _f32(g_out_pos) if g_out_pos is not None else None
    check(lib().fs2k_lr_bwd(_p(g1), _p(g2), _p(cum), B, T, D, F_out, _p(dx), _stream()), "fs2k_lr_bwd")
    _count()
    return dx


def scatter_add_rows(dtable, g1, ids, g2=None, skip_id: int = -1):
    g1 = _f32(g1, "g1")
    if g2 is not None:
        g2 = _f32(g2, "g2")
    ids = ids.contiguous()
    D = dtable.shape[1]
    check(lib().fs2k_scatter_add_rows(_p(g1), _p(g2), _p(ids), int(ids.dtype == torch.int64), ids.numel(), D, int(skip_id), _p(dtable), _stream()), "fs2k_scatter_add_rows")
    _count()
    return dtable


def rows_sum_scatter(drows, g, ids):
    g = _f32(g, "g")
    B, L, D = g.shape
    check(lib().fs2k_rows_sum_scatter(_p(g), _p(_i32(ids)) if ids is not None else None, B, L, D, _p(drows), _stream()), "fs2k_rows_sum_scatter")
    _count()
    return drows


def aligner_bwd(g_soft, g_logprob, soft, logprob, prior, key_lens, q, k):
    B, F, C = q.shape
    T = k.shape[1]
    dd = torch.empty((B, F, T), dtype=torch.float32, device=q.device)
    dq, dk = torch.empty_like(q), torch.empty_like(k)
    gs = _f32(g_soft) if g_soft is not None else None
    gl = _f32(g_logprob) if g_logprob is not None else None
    kl = _i32(key_lens) if key_lens is not None else None
    check(lib().fs2k_aligner_bwd(_p(gs), _p(gl), _p(soft), _p(logprob), _p(prior), _p(kl), _p(q), _p(k), B, F, T, C, _p(dd), _p(dq), _p(dk), _stream()), "fs2k_aligner_bwd")
    _count(3)
    return dq, dk
