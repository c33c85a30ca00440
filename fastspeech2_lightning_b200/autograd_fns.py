"""`torch.autograd.Function`s over libfs2k kernels (forward AND backward are kernel launches).

`autograd.py` routes here whenever a gradient is required; ops without trainable inputs are plain
forwards.  Backward kernels live in csrc/*_bwd.cu (built in the training milestone); until an op's
backward exists its Function raises in `backward`, never silently falls back to torch math.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

DROPOUT_IMPLEMENTED = False


def _grad_on(*ts) -> bool:
    return torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in ts)


class _Pending(torch.autograd.Function):
    """Marks an output as depending on its inputs; raises if a backward is actually requested."""

    @staticmethod
    def forward(ctx, name, out, *inputs):
        ctx.name = name
        return out.view_as(out)

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError(f"backward kernel for '{ctx.name}' is not built yet")


def _pending(name, out, *inputs):
    if _grad_on(*inputs):
        if isinstance(out, tuple):
            return tuple(_Pending.apply(name, o, *inputs) if o is not None and o.is_floating_point() else o for o in out)
        return _Pending.apply(name, out, *inputs)
    return out


# ---------------------------------------------------------------------------------------------
def layernorm(x, weight, bias, eps, dropout=0.0):
    assert not dropout, "dropout kernels pending"
    return _pending("layernorm", ops.layernorm(x, weight.detach(), bias.detach(), eps), x, weight, bias)


def linear(x, weight, bias, act, alpha, residual, dropout=0.0):
    assert not dropout, "dropout kernels pending"
    y = ops.gemm(x.detach(), weight.detach(), None if bias is None else bias.detach(), act=act, alpha=alpha,
                 residual=None if residual is None else residual.detach())
    return _pending("linear", y, x, weight, bias, residual)


def conv1d(x, weight, bias, act):
    w = ops.conv_weight_taps(weight)
    y = ops.gemm(x.detach(), w, None if bias is None else bias.detach(), taps_pad=(weight.shape[-1] - 1) // 2, act=act)
    return _pending("conv1d", y, x, weight, bias)


def conv1d_bn_act(x, weight, bias, bn, act, training, dropout=0.0):
    assert not dropout, "dropout kernels pending"
    w = ops.conv_weight_taps(weight)
    pad = (weight.shape[-1] - 1) // 2
    xd, bd = x.detach(), None if bias is None else bias.detach()
    if not training:
        scale, shift = ops.bn_scale_shift(bn, None, False)
        y = ops.gemm(xd, w, bd, taps_pad=pad, scale=scale, shift=shift, act=act)
    else:
        z = ops.gemm(xd, w, bd, taps_pad=pad)
        scale, shift = ops.bn_scale_shift(bn, z, True)
        y = ops.affine_act(z, scale, shift, act)
    return _pending("conv1d_bn_act", y, x, weight, bias, bn.weight, bn.bias)


def glu_dwconv_bn_silu(h, dw_weight, dw_bias, bn, training):
    C = dw_weight.shape[0]
    hd, wd, bd = h.detach(), dw_weight.detach(), dw_bias.detach()
    if not training:
        scale, shift = ops.bn_scale_shift(bn, None, False)
        y = ops.dwconv(hd, wd, bd, channels=C, glu=True, scale=scale, shift=shift)
    else:
        z = ops.dwconv(hd, wd, bd, channels=C, glu=True)
        scale, shift = ops.bn_scale_shift(bn, z, True)
        y = ops.affine_act(z, scale, shift, "silu")
    return _pending("glu_dwconv_bn_silu", y, h, dw_weight, dw_bias, bn.weight, bn.bias)


def dwconv(x, weight, bias):
    y = ops.dwconv(x.detach(), weight.detach(), bias.detach(), channels=weight.shape[0])
    return _pending("dwconv", y, x, weight, bias)


def attention(qkv, lengths, heads, dropout=0.0):
    assert not dropout, "dropout kernels pending"
    return _pending("attention", ops.attention(qkv.detach(), lengths, heads), qkv)


def rowdot(x, weight, bias, mask):
    y = ops.rowdot(x.detach(), weight.detach().reshape(-1), bias.detach(), mask)
    return _pending("rowdot", y, x, weight, bias)


def aligner_scores(q, k, prior, key_lens):
    soft, logprob = ops.aligner_scores(q.detach(), k.detach(), prior, key_lens)
    return _pending("aligner_scores", (soft, logprob), q, k)


def length_regulate(x, cum, total, width, inv_freq):
    out, out_pos, mask, _ = ops.lr_gather(x.detach(), cum, total, width, inv_freq, want_out=True)
    out, out_pos = _pending("length_regulate", (out, out_pos), x)
    return out, out_pos, mask


def bucketize_embed_add(v, scale, bins, table, x, return_scaled=False):
    y, ids, vs = ops.bucketize_embed_add(v.detach().contiguous(), bins.detach(), table.detach(), x.detach(), scale=scale,
                                         want_ids=True, want_scaled=return_scaled)
    y = _pending("bucketize_embed_add", y, x, table)
    if return_scaled:
        return y, ids, _pending("scale", vs, v)
    return y, ids


def embedding_lookup(ids, table):
    return _pending("embedding_lookup", ops.gather_rows(table.detach(), ids), table)


def scale(x, factor):
    return _pending("scale", ops.axpby(x.detach(), float(factor)), x)


def embed_posenc(text, table, inv_freq, lens, padding_idx):
    emb, x = ops.embed_posenc(text.contiguous(), table.detach(), inv_freq, lens)
    return _pending("embed_posenc", (emb, x), table)


def add_posenc(x, inv_freq, lens):
    return _pending("add_posenc", ops.add_posenc(x.detach(), inv_freq, lens), x)


def add_rows(x, rows):
    y = ops.add_rows(x.detach(), [(r.detach().contiguous(), i) for r, i in rows])
    return _pending("add_rows", y, x, *[r for r, _ in rows])


def add(a, b):
    return _pending("add", ops.axpby(a.detach(), 1.0, b.detach(), 1.0), a, b)


def masked_loss(pred, target, mask, kind, weight, log1p_int_target=False):
    loss = ops.masked_loss_fwd(pred.detach(), target.detach(), mask, kind, weight, log1p_int_target)
    return _pending("masked_loss", loss, pred)


def attn_bin_loss(hard, soft, eps=1e-12):
    loss, _ = ops.bin_loss_fwd(hard, soft.detach(), eps)
    return _pending("attn_bin_loss", loss, soft)


def tanh_row(table, index):
    return _pending("tanh_row", ops.tanh(table.detach()[index : index + 1].contiguous()), table)
