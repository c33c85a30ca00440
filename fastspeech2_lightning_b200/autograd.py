"""Differentiable building blocks: every function runs libfs2k kernels forward and — when
gradients are required — registers a `torch.autograd.Function` whose backward also runs
libfs2k kernels (see autograd_fns.py).  Without grad mode these are plain kernel launches.
"""
from __future__ import annotations


import torch

from . import ops


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors)


def _no_dropout(p: float):
    if p and p > 0.0:
        from . import autograd_fns

        if not autograd_fns.DROPOUT_IMPLEMENTED:
            raise NotImplementedError("training-mode dropout > 0 is not built yet; run with dropout 0 or eval()")


def layernorm(x, weight, bias, eps: float = 1e-5, dropout: float = 0.0):
    if _needs_grad(x, weight, bias) or dropout > 0.0:
        from . import autograd_fns as fns

        return fns.layernorm(x, weight, bias, eps, dropout)
    return ops.layernorm(x, weight, bias, eps)


def layernorm_fork(x, weight, bias, eps: float = 1e-5):
    """(LayerNorm(x), skip) for a pre-norm residual block: use `skip` (== x) as the block's residual input, so that the
    two gradients of x meet in one backward launch (autograd_fns._LayerNormFork)."""
    if _needs_grad(x, weight, bias):
        from . import autograd_fns as fns

        return fns.layernorm_fork(x, weight, bias, eps)
    return ops.layernorm(x, weight, bias, eps), x


def linear(x, weight, bias=None, act=None, alpha: float = 1.0, residual=None, dropout: float = 0.0):
    """act(x·Wᵀ + b)·alpha (+ residual); weight [N,K]."""
    if _needs_grad(x, weight, bias, residual) or dropout > 0.0:
        from . import autograd_fns as fns

        return fns.linear(x, weight, bias, act, alpha, residual, dropout)
    return ops.gemm(x, weight.detach(), bias, act=act, alpha=alpha, residual=residual)


def conv1d(x, weight, bias, act=None):
    """nn.Conv1d(k, padding=(k-1)//2) on channels-last x [B,L,Cin]; weight [Cout,Cin,k]."""
    if _needs_grad(x, weight, bias):
        from . import autograd_fns as fns

        return fns.conv1d(x, weight, bias, act)
    w = ops.conv_weight_taps(weight)
    return ops.gemm(x, w, bias, taps_pad=(weight.shape[-1] - 1) // 2, act=act)


def conv1d_bn_act(x, weight, bias, bn, act, training: bool, dropout: float = 0.0):
    """Conv1d(k) → BatchNorm1d → act (→ dropout): one PostNet block (fs2/layers.py:157-212)."""
    if _needs_grad(x, weight, bias, bn.weight) or dropout > 0.0:
        from . import autograd_fns as fns

        return fns.conv1d_bn_act(x, weight, bias, bn, act, training, dropout)
    w = ops.conv_weight_taps(weight)
    pad = (weight.shape[-1] - 1) // 2
    if not training:
        scale, shift = ops.bn_scale_shift(bn, None, False)
        return ops.gemm(x, w, bias, taps_pad=pad, scale=scale, shift=shift, act=act)
    z = ops.gemm(x, w, bias, taps_pad=pad)
    scale, shift = ops.bn_scale_shift(bn, z, True)
    return ops.affine_act(z, scale, shift, act)


def glu_dwconv_bn_silu(h, dw_weight, dw_bias, bn, training: bool):
    """GLU → depthwise conv → BatchNorm1d → SiLU (torchaudio conformer.py:50-65); h [B,L,2C] → [B,L,C]."""
    C = dw_weight.shape[0]
    if _needs_grad(h, dw_weight, dw_bias, bn.weight):
        from . import autograd_fns as fns

        return fns.glu_dwconv_bn_silu(h, dw_weight, dw_bias, bn, training)
    if not training:
        scale, shift = ops.bn_scale_shift(bn, None, False)
        return ops.dwconv(h, dw_weight, dw_bias, channels=C, glu=True, scale=scale, shift=shift)
    z = ops.dwconv(h, dw_weight, dw_bias, channels=C, glu=True)
    scale, shift = ops.bn_scale_shift(bn, z, True)
    return ops.affine_act(z, scale, shift, "silu")


def dwconv(x, weight, bias):
    if _needs_grad(x, weight, bias):
        from . import autograd_fns as fns

        return fns.dwconv(x, weight, bias)
    return ops.dwconv(x, weight, bias, channels=weight.shape[0])


def attention(qkv, lengths, heads: int, dropout: float = 0.0, order=None):
    if _needs_grad(qkv) or dropout > 0.0:
        from . import autograd_fns as fns

        return fns.attention(qkv, lengths, heads, dropout, order)
    return ops.attention(qkv, lengths, heads, order=order)


def qkv_attention(x, in_proj_weight, in_proj_bias, lengths, heads: int, dropout: float = 0.0, order=None):
    """MultiheadAttention's in-projection + masked scaled-dot-product attention (torchaudio conformer.py:193-202).
    bf16 mode (head_dim 128): one bf16 qkv tensor from the GEMM epilogue feeds the tcgen05 attention kernels;
    other modes: the fp32 GEMM followed by the fp32 attention kernel."""
    D = x.shape[-1]
    if (ops.PRECISION == "bf16" and D // heads == 128 and lib_supports_bf16_linear(D, in_proj_weight.shape[0])):
        if _needs_grad(x, in_proj_weight, in_proj_bias) or dropout > 0.0:
            from . import autograd_fns as fns

            return fns.qkv_attention_bf16(x, in_proj_weight, in_proj_bias, lengths, heads, dropout, order)
        w16, _ = ops.bf16_weight(in_proj_weight.detach())
        _, qkv16, _ = ops.gemm_bf16(x, w16, in_proj_bias, want_c=False, want_c16=True)
        return ops.attention_bf16(qkv16, lengths, heads, order=order)
    qkv = linear(x, in_proj_weight, in_proj_bias)
    return attention(qkv, lengths, heads, dropout=dropout, order=order)


def lib_supports_bf16_linear(K: int, N: int) -> bool:
    from ._lib import lib

    return bool(lib().fs2k_gemm_bf16_supported(K, N, K, 1, 0, 0))


def rowdot(x, weight, bias, mask):
    """nn.Linear(D→1) + squeeze(-1) + ·mask (fs2/variance_adaptor.py:58-61)."""
    if _needs_grad(x, weight, bias):
        from . import autograd_fns as fns

        return fns.rowdot(x, weight, bias, mask)
    return ops.rowdot(x, weight.reshape(-1), bias, mask)


def aligner_scores(q, k, prior, key_lens):
    if _needs_grad(q, k):
        from . import autograd_fns as fns

        return fns.aligner_scores(q, k, prior, key_lens)
    return ops.aligner_scores(q, k, prior, key_lens)
