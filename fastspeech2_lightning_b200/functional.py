"""Compositions of libfs2k kernels that implement the reference's blocks.

Each function takes the parameter-holding `nn.Module` (whose state-dict layout mirrors the
reference) plus channels-last CUDA activations, and chains `ops.*` kernel launches.  Citations
are to the reference files (under /root/reference) or torchaudio's `models/conformer.py`.
"""
from __future__ import annotations


import torch

from . import autograd as ag


# ---------------------------------------------------------------------------------------------
# Conformer  (torchaudio conformer.py:176-212, :273-293)
# ---------------------------------------------------------------------------------------------
def _ffn_half(x, ffn, training: bool, p_drop: float):
    """x + ½·FFN(x):  LN → Linear → SiLU → Dropout → Linear → Dropout (conformer.py:103-108,:185-187)."""
    from . import ops

    s = ffn.sequential
    p = p_drop if training else 0.0
    D, Hd = s[1].weight.shape[1], s[1].weight.shape[0]
    if ops.PRECISION == "bf16" and ops.bf16_dact_ok(D, Hd) and Hd % 8 == 0 and D % 16 == 0 and s[1].bias is not None and s[4].bias is not None:
        if ag._needs_grad(x, *ffn.parameters()) or p > 0.0:
            from . import autograd_fns as fns

            return fns.ffn_half_bf16(x, ffn, p)
        # synthesis: the hidden activation only exists in bf16 between the two GEMMs
        ln = ops.layernorm(x, s[0].weight, s[0].bias, s[0].eps)
        w1_16, _ = ops.bf16_weight(s[1].weight.detach())
        w2_16, _ = ops.bf16_weight(s[4].weight.detach())
        _, h16, _ = ops.gemm_bf16(ln, w1_16, s[1].bias, act="silu", want_c=False, want_c16=True)
        return ops.gemm_bf16(h16, w2_16, s[4].bias, alpha=0.5, residual=x)[0]
    ln, skip = ag.layernorm_fork(x, s[0].weight, s[0].bias, s[0].eps)
    h = ag.linear(ln, s[1].weight, s[1].bias, act="silu", dropout=p)
    return ag.linear(h, s[4].weight, s[4].bias, alpha=0.5, residual=skip, dropout=p)


def conformer_layer(x, lengths, layer, training: bool, order=None):
    p = layer.dropout_p
    x = _ffn_half(x, layer.ffn1, training, p)
    # self-attention block (:191-203)
    ln, skip = ag.layernorm_fork(x, layer.self_attn_layer_norm.weight, layer.self_attn_layer_norm.bias, layer.self_attn_layer_norm.eps)
    mha = layer.self_attn
    o = ag.qkv_attention(ln, mha.in_proj_weight, mha.in_proj_bias, lengths, layer.num_heads,
                         dropout=mha.dropout if training else 0.0, order=order)
    x = ag.linear(o, mha.out_proj.weight, mha.out_proj.bias, residual=skip, dropout=p if training else 0.0)
    # convolution module (:42-75, :168-174)
    cm = layer.conv_module
    seq = cm.sequential
    ln, skip = ag.layernorm_fork(x, cm.layer_norm.weight, cm.layer_norm.bias, cm.layer_norm.eps)
    h = ag.linear(ln, seq[0].weight.squeeze(-1), seq[0].bias)  # pointwise D → 2D (pre-GLU)
    d = ag.glu_dwconv_bn_silu(h, seq[2].weight, seq[2].bias, seq[3], training)
    x = ag.linear(d, seq[5].weight.squeeze(-1), seq[5].bias, residual=skip, dropout=p if training else 0.0)
    x = _ffn_half(x, layer.ffn2, training, p)
    fl = layer.final_layer_norm
    return ag.layernorm(x, fl.weight, fl.bias, fl.eps)


def conformer_stack(x, lengths, conformer, training: bool):
    lengths = lengths.to(torch.int32)
    from . import ops

    order = ops.attention_order(lengths)  # once per stack: the attention CTAs of every layer run longest utterance first
    for layer in conformer.conformer_layers:
        x = conformer_layer(x, lengths, layer, training, order)
    return x


# ---------------------------------------------------------------------------------------------
# variance predictor  (fs2/variance_adaptor.py:18-62, fs2/layers.py:20-48, fs2/blocks.py:4-19)
# ---------------------------------------------------------------------------------------------
def variance_predictor(x, mask, vp, training: bool):
    for layer in vp.conv:
        conv = layer.layers[0].module
        ln = layer.layers[2]
        p = layer.layers[3].p if training else 0.0
        if hasattr(conv, "model"):  # DepthwiseSeparableConv1d: depthwise k → pointwise 1×1
            h = ag.dwconv(x, conv.model[0].weight, conv.model[0].bias)
            h = ag.linear(h, conv.model[1].weight.squeeze(-1), conv.model[1].bias, act="relu")
        else:  # plain Conv1d k
            h = ag.conv1d(x, conv.weight, conv.bias, act="relu")
        x = ag.layernorm(h, ln.weight, ln.bias, ln.eps, dropout=p)
    return ag.rowdot(x, vp.linear.weight, vp.linear.bias, mask)


# ---------------------------------------------------------------------------------------------
# PostNet  (fs2/layers.py:143-212)
# ---------------------------------------------------------------------------------------------
def postnet(x, pn, training: bool):
    from . import ops

    n = len(pn.convolutions)
    p_drop = 0.5 if training and pn.dropout_in_training else 0.0
    if ops.PRECISION == "bf16" and _postnet_bf16_ok(pn, x):
        if ag._needs_grad(x, *pn.parameters()) or p_drop > 0.0 or training:
            from . import autograd_fns as fns

            return fns.postnet_bf16(x, pn, training, p_drop)
        # synthesis: BatchNorm folded into the convolution's epilogue, bf16 activations between the layers
        h = x.detach()
        for i, block in enumerate(pn.convolutions):
            conv, bn = block[0].conv, block[1]
            w_taps = ops.conv_weight_taps(conv.weight)
            taps, N, K = w_taps.shape
            w16, _ = ops.bf16_weight(w_taps)
            scale, shift = ops.bn_scale_shift(bn, None, False)
            last = i == n - 1
            c, c16, _ = ops.gemm_bf16(h, w16, conv.bias, taps_pad=(taps - 1) // 2, scale=scale, shift=shift, act=None if last else "tanh",
                                      want_c=last, want_c16=not last, block_n_hint=256 if (taps * K >= 1024 and N % 256 == 0) else 0)
            h = c if last else c16
        return h
    for i, block in enumerate(pn.convolutions):
        conv, bn = block[0].conv, block[1]
        act = "tanh" if i < n - 1 else None
        x = ag.conv1d_bn_act(x, conv.weight, conv.bias, bn, act, training, dropout=p_drop)
    return x


def _postnet_bf16_ok(pn, x) -> bool:
    """Every layer's shapes are taken by the TMA-fed bf16 kernels (channels multiples of 16: 80 / 512 in the base config)."""
    if not x.is_cuda or x.dim() != 3:
        return False
    for block in pn.convolutions:
        w = block[0].conv.weight
        if w.shape[0] % 16 or w.shape[1] % 16 or (w.shape[0] > 256 and w.shape[0] % 128):
            return False
    return True


# ---------------------------------------------------------------------------------------------
# aligner  (fs2/attn/attention.py:195-251)
# ---------------------------------------------------------------------------------------------
def conv_attention(att, queries_blc, keys_blc, key_lens, prior):
    """queries [B,F,n_mel], keys [B,T,n_text] channels-last → (attn_soft, attn_logprob) [B,1,F,T]."""
    kp, qp = att.key_proj, att.query_proj
    k = ag.conv1d(keys_blc, kp[0].conv.weight, kp[0].conv.bias, act="relu")
    k = ag.conv1d(k, kp[2].conv.weight, kp[2].conv.bias)
    q = ag.conv1d(queries_blc, qp[0].conv.weight, qp[0].conv.bias, act="relu")
    q = ag.conv1d(q, qp[2].conv.weight, qp[2].conv.bias, act="relu")
    q = ag.conv1d(q, qp[4].conv.weight, qp[4].conv.bias)
    return ag.aligner_scores(q, k, prior, key_lens)
