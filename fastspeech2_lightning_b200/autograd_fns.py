"""`torch.autograd.Function`s over libfs2k kernels: forward AND backward are kernel launches;
torch.autograd only sequences them.  `autograd.py` routes here whenever a gradient is required.
"""
from __future__ import annotations


import torch
from torch.autograd.function import once_differentiable

from . import ops

DROPOUT_IMPLEMENTED = True


def _d(t):
    return None if t is None else t.detach()


def _seed(p) -> int:
    """A fresh dropout seed from torch's CPU generator (no device sync); 0 when dropout is off."""
    return int(torch.randint(0, 2**31 - 1, (1,)).item()) if p else 0


# ---------------------------------------------------------------------------------------------
# dropout (Philox-free counter hash: the mask is regenerated from (seed, offset) in backward)
# ---------------------------------------------------------------------------------------------
class _Dropout(torch.autograd.Function):
    """dropout(x) (+ residual) in one launch."""

    @staticmethod
    def forward(ctx, x, residual, p):
        seed = _seed(p)
        ctx.seed, ctx.p, ctx.has_res = seed, p, residual is not None
        return ops.dropout(x, p, seed, residual)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = g.contiguous()
        return ops.dropout(g, ctx.p, ctx.seed), (g if ctx.has_res else None), None


def dropout(x, p: float, residual=None):
    if not p:
        return x if residual is None else add(x, residual)
    return _Dropout.apply(x.contiguous(), residual, float(p))


# ---------------------------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, p):
        seed = _seed(p)
        y, mean, rstd = ops.layernorm(x, weight, bias, eps, save_stats=True, dropout_p=p, seed=seed)
        ctx.save_for_backward(x, weight, mean, rstd)
        ctx.drop = (p, seed)
        ctx.params = (weight, bias)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, weight, mean, rstd = ctx.saved_tensors
        direct = _direct_grads(*ctx.params)
        dx, dw, db = ops.layernorm_bwd(g.contiguous(), x, mean, rstd, weight, *ctx.drop, accumulate_into=direct)
        if direct is not None:
            dw = db = None
        return dx, dw, db, None, None


def layernorm(x, weight, bias, eps, dropout_p=0.0):
    """LayerNorm with the following Dropout fused into the same launch (forward and backward)."""
    return _LayerNorm.apply(x.contiguous(), weight, bias, eps, float(dropout_p))


class _LayerNormFork(torch.autograd.Function):
    """x → (LayerNorm(x), x): the pre-norm residual block `x + f(LayerNorm(x))` takes its skip connection from the second
    output, so both gradients of x arrive at ONE node and the backward adds them inside the LayerNorm-backward launch
    (`fs2k_layernorm_bwd_add`) instead of in a separate elementwise add per residual block."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        y, mean, rstd = ops.layernorm(x, weight, bias, eps, save_stats=True)
        ctx.save_for_backward(x, weight, mean, rstd)
        ctx.params = (weight, bias)
        return y, x.view_as(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, g, g_skip):
        x, weight, mean, rstd = ctx.saved_tensors
        direct = _direct_grads(*ctx.params)
        if g is None:  # only the skip path was used
            return g_skip, None, None, None
        dx, dw, db = ops.layernorm_bwd(g.contiguous(), x, mean, rstd, weight, accumulate_into=direct,
                                       add=None if g_skip is None else g_skip.contiguous())
        if direct is not None:
            dw = db = None
        return dx, dw, db, None


def layernorm_fork(x, weight, bias, eps):
    return _LayerNormFork.apply(x.contiguous(), weight, bias, eps)


# ---------------------------------------------------------------------------------------------
class WgradSink:
    """Weight/bias gradients off the critical path.  The data-gradient chain (dgrad GEMM → LayerNorm backward → …) is
    what the next backward node waits for; a weight gradient is a leaf.  With a sink installed every contraction's
    wgrad + bias column-sum runs on a second stream and accumulates straight into `param.grad` (FusedAdamW's flat
    gradient buffer), autograd receives None for those inputs (no AccumulateGrad launch), and `join()` — called once
    before the optimizer — makes the main stream wait for the side stream.  Under CUDA-graph capture the two streams
    become parallel branches of the graph.  Gradient tensors read by the side stream are kept referenced until the
    join so the allocator cannot hand their memory to a later main-stream kernel."""

    def __init__(self):
        self.stream = torch.cuda.Stream()
        self.keep = []

    def join(self):
        torch.cuda.current_stream().wait_stream(self.stream)
        self.keep.clear()

    def fence(self):
        """Main stream waits for what the side stream was given during the forward (transposed weights, CTC loss)."""
        torch.cuda.current_stream().wait_stream(self.stream)

    def prefetch_transposed(self, w_taps):
        """The dgrad GEMM's operand ([taps,N,K] → reversed-tap [taps,K,N]) depends on the weights only: re-pack it on
        the side stream during the forward instead of on the critical path of the backward."""
        self.stream.wait_stream(torch.cuda.current_stream())  # w_taps itself may just have been re-packed there
        with torch.cuda.stream(self.stream):
            return ops.weight_taps_transposed(w_taps)


_SINK = None


def set_wgrad_sink(sink):
    """Install (or remove with None) the side-stream weight-gradient sink; returns the previous one."""
    global _SINK
    prev, _SINK = _SINK, sink
    return prev


def _direct_grads(*params):
    """The parameters' flat-gradient views when a sink is installed and every one of them is a leaf with a .grad —
    the backward kernels then add onto them and autograd receives None (no AccumulateGrad launch)."""
    if _SINK is None or any(p is None or not p.is_leaf or not p.requires_grad or p.grad is None for p in params):
        return None
    return [p.grad for p in params]


def _prefetch_wt(w_taps, needs_dx):
    """Transposed weights for the data-gradient GEMM, re-packed on the side stream during the forward — not needed in
    the bf16 mode, whose dgrad reads the forward weights as an MN-major operand."""
    if _SINK is None or not needs_dx:
        return None
    if ops.PRECISION in ("bf16", "bf16x3") and ops.bf16_dgrad_ok(w_taps):
        return None
    return _SINK.prefetch_transposed(w_taps)


def _gemm_backward(g, x, w_taps, pad, conv_layout, needs_x, needs_w, needs_b, weight=None, bias=None, wt=None, precision=None):
    """Shared by every contraction: g = gradient w.r.t. the pre-activation conv output.  `precision` = the arithmetic
    mode the forward ran in (the backward of a full-precision region stays full precision in the bf16 mode)."""
    with ops.precision_scope(precision):
        return _gemm_backward_impl(g, x, w_taps, pad, conv_layout, needs_x, needs_w, needs_b, weight, bias, wt)


def _gemm_backward_impl(g, x, w_taps, pad, conv_layout, needs_x, needs_w, needs_b, weight, bias, wt):
    taps, N, K = w_taps.shape
    dx = dw = db = None
    sink = _SINK
    if sink is not None and (needs_w or needs_b) and weight is not None and weight.is_leaf and weight.grad is not None and (
            not needs_b or (bias is not None and bias.is_leaf and bias.grad is not None)):
        sink.stream.wait_stream(torch.cuda.current_stream())  # g and x are complete on the main stream
        sink.keep.append((g, x))
        with torch.cuda.stream(sink.stream), ops.backward_precision():
            if needs_b:
                ops.colsum(g, out=bias.grad, accumulate=True)
            if needs_w:
                ops.gemm_wgrad(g, x, taps, pad, conv_layout, accumulate_into=weight.grad)
        needs_w = needs_b = False
    if needs_b:
        db = ops.colsum(g)
    with ops.backward_precision():
        if needs_x:
            dx = ops.gemm_dgrad(g, w_taps, pad, wt)
        if needs_w:
            dw = ops.gemm_wgrad(g, x, taps, pad, conv_layout)
    return dx, dw, db


class _Gemm(torch.autograd.Function):
    """y = act(conv(x, W) + b)·alpha + residual   (Linear when W is 2-D)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, act, alpha, p=0.0):
        conv_layout = weight.dim() == 3
        w_taps = ops.conv_weight_taps(weight) if conv_layout else weight.reshape(1, *weight.shape)
        pad = (w_taps.shape[0] - 1) // 2
        aux = None
        seed = _seed(p)
        ctx.drop = (p, seed)
        if act == "silu":  # keep the pre-activation for silu'; SiLU (+ Dropout) fused (bf16 mode: into the GEMM epilogue)
            aux, y = ops.gemm_silu_pair(x, w_taps, bias, pad, residual, p, seed)
            if alpha != 1.0:
                raise NotImplementedError("alpha with silu")
        elif p:
            # residual + Dropout(Linear(x))·alpha in the GEMM epilogue; the backward regenerates the mask in act_bwd
            if act is not None:
                raise NotImplementedError("fused dropout is wired for the SiLU and the plain epilogue")
            y = ops.gemm(x, w_taps, bias, taps_pad=pad, alpha=alpha, residual=residual, dropout_p=p, seed=seed)
        else:
            y = ops.gemm(x, w_taps, bias, taps_pad=pad, act=act, alpha=alpha, residual=residual)
            if act in ("relu", "tanh"):
                if residual is not None or alpha != 1.0:
                    raise NotImplementedError("relu/tanh epilogue with residual/alpha in training")
                aux = y
        ctx.save_for_backward(x, w_taps, aux)
        ctx.meta = (act, alpha, pad, conv_layout, bias is not None, residual is not None)
        ctx.params = (weight, bias)
        ctx.precision = ops.PRECISION
        ctx.wt = _prefetch_wt(w_taps, ctx.needs_input_grad[0])
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w_taps, aux = ctx.saved_tensors
        act, alpha, pad, conv_layout, has_bias, has_res = ctx.meta
        g = g.contiguous()
        gz = g if (act is None and alpha == 1.0 and not ctx.drop[0]) else ops.act_bwd(g, aux, act, alpha, None, *ctx.drop)
        dx, dw, db = _gemm_backward(gz, x, w_taps, pad, conv_layout, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                    has_bias and ctx.needs_input_grad[2], *ctx.params, wt=ctx.wt, precision=ctx.precision)
        return dx, dw, db, (g if has_res and ctx.needs_input_grad[3] else None), None, None, None


def linear(x, weight, bias, act, alpha, residual, dropout_p=0.0):
    if dropout_p and act == "silu" and residual is None:
        return _Gemm.apply(x.contiguous(), weight, bias, None, act, alpha, float(dropout_p))
    if dropout_p and act is None:
        # reference order: residual + Dropout(Linear(...))·alpha — all of it in the GEMM epilogue
        return _Gemm.apply(x.contiguous(), weight, bias, residual, None, alpha, float(dropout_p))
    if dropout_p:
        y = _Gemm.apply(x.contiguous(), weight, bias, None, act, alpha, 0.0)
        return dropout(y, dropout_p, residual)
    return _Gemm.apply(x.contiguous(), weight, bias, residual, act, alpha, 0.0)


def conv1d(x, weight, bias, act):
    return _Gemm.apply(x.contiguous(), weight, bias, None, act, 1.0, 0.0)


# ---------------------------------------------------------------------------------------------
class _ConvBnAct(torch.autograd.Function):
    """One PostNet block: Conv1d(k) → BatchNorm1d → act (fs2/layers.py:157-212)."""

    @staticmethod
    def forward(ctx, x, weight, bias, bn_w, bn_b, bn, act, training, p=0.0):
        w_taps = ops.conv_weight_taps(weight)
        pad = (w_taps.shape[0] - 1) // 2
        z = ops.gemm(x, w_taps, bias, taps_pad=pad)
        scale, shift, mean, rstd = ops.bn_scale_shift(bn, z, training, save_stats=True)
        seed = _seed(p)
        y = ops.affine_act(z, scale, shift, act, None, p, seed)  # BatchNorm affine + act + Dropout in one launch
        ctx.save_for_backward(x, w_taps, z, scale, shift, mean, rstd)
        ctx.meta = (act, training, pad, bias is not None, p, seed)
        ctx.params = (weight, bias)
        ctx.bn_params = (bn_w, bn_b)
        ctx.precision = ops.PRECISION
        ctx.wt = _prefetch_wt(w_taps, ctx.needs_input_grad[0])
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w_taps, z, scale, shift, mean, rstd = ctx.saved_tensors
        act, training, pad, has_bias, p, seed = ctx.meta
        direct = _direct_grads(*ctx.bn_params)
        gz, dgamma, dbeta = ops.bn_act_bwd(g.contiguous(), z, scale, shift, mean, rstd, act, training, p, seed, accumulate_into=direct)
        if direct is not None:
            dgamma = dbeta = None
        dx, dw, db = _gemm_backward(gz, x, w_taps, pad, True, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                    has_bias and ctx.needs_input_grad[2], *ctx.params, wt=ctx.wt, precision=ctx.precision)
        return dx, dw, db, dgamma, dbeta, None, None, None, None


def conv1d_bn_act(x, weight, bias, bn, act, training, dropout_p=0.0):
    return _ConvBnAct.apply(x.contiguous(), weight, bias, bn.weight, bn.bias, bn, act, training, float(dropout_p))


class _FfnHalfBf16(torch.autograd.Function):
    """x + ½·Dropout(W2·Dropout(SiLU(W1·LN(x) + b1)) + b2) — one half-step feed-forward module of a Conformer layer
    (torchaudio conformer.py:91-119,185-187,207-209) in the bf16 mode as ONE autograd node.

    forward : LN → row-panel GEMM whose epilogue writes the pre-activation AND dropout(SiLU(·)) in bf16 (the 1024-wide hidden
              tensors never exist in fp32) → TMA-fed GEMM with bias, dropout, ½ and the residual in its epilogue.
    backward: dropout-mask/½ on g → row-panel data-gradient GEMM whose epilogue multiplies by SiLU'(pre)·mask and writes bf16 →
              TMA-fed data-gradient GEMM → LayerNorm backward with the skip connection's gradient added in the same launch;
              weight / bias gradients (bf16 operands by TMA) on the side stream."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, eps, w1, b1, w2, b2, p_drop):
        ln, mean, rstd = ops.layernorm(x, ln_w, ln_b, eps, save_stats=True)
        seed1, seed2 = _seed(p_drop), _seed(p_drop)
        w1_16, _ = ops.bf16_weight(w1.detach())
        w2_16, _ = ops.bf16_weight(w2.detach())
        _, h16, pre16 = ops.gemm_bf16(ln, w1_16, b1.detach(), act="silu", want_c=False, want_c16=True, want_pre="bf16",
                                      dropout_p=p_drop, seed=seed1)
        y, _, _ = ops.gemm_bf16(h16, w2_16, b2.detach(), alpha=0.5, residual=x, dropout_p=p_drop, seed=seed2)
        ctx.save_for_backward(x, ln, mean, rstd, pre16, h16, w1_16, w2_16)
        ctx.cfg = (p_drop, seed1, seed2)
        ctx.params = (ln_w, ln_b, w1, b1, w2, b2)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, ln, mean, rstd, pre16, h16, w1_16, w2_16 = ctx.saved_tensors
        p_drop, seed1, seed2 = ctx.cfg
        ln_w, ln_b, w1, b1, w2, b2 = ctx.params
        g = g.contiguous()
        gz = ops.act_bwd(g, None, None, 0.5, None, p_drop, seed2)                       # ½ and the outer dropout mask
        _, gz1_16 = ops.gemm_bf16_dact(gz, w2_16, pre16, "silu", w_mn=True, dropout_p=p_drop, seed=seed1)   # (gz·W2)∘SiLU'(pre)∘mask
        dln, _, _ = ops.gemm_bf16(gz1_16, w1_16.reshape(1, *w1_16.shape), None, w_mn=True)
        sink = _SINK
        grads = [None] * 6
        direct = _direct_grads(w1, b1, w2, b2)
        if sink is not None and direct is not None:
            sink.stream.wait_stream(torch.cuda.current_stream())
            sink.keep.append((gz, gz1_16, h16, ln))
            with torch.cuda.stream(sink.stream):
                ops.colsum(gz, out=b2.grad, accumulate=True)
                ops.gemm_wgrad_bf16(gz, h16, 1, 0, False, accumulate_into=w2.grad)
                ops.colsum(gz1_16, out=b1.grad, accumulate=True)
                ops.gemm_wgrad_bf16(gz1_16, ln, 1, 0, False, accumulate_into=w1.grad)
        else:
            grads[5], grads[4] = ops.colsum(gz), ops.gemm_wgrad_bf16(gz, h16, 1, 0, False)
            grads[3], grads[2] = ops.colsum(gz1_16), ops.gemm_wgrad_bf16(gz1_16, ln, 1, 0, False)
        direct_ln = _direct_grads(ln_w, ln_b)
        dx, dgamma, dbeta = ops.layernorm_bwd(dln, x, mean, rstd, ln_w, accumulate_into=direct_ln, add=g)
        if direct_ln is None:
            grads[0], grads[1] = dgamma, dbeta
        return dx, grads[0], grads[1], None, grads[2], grads[3], grads[4], grads[5], None


def ffn_half_bf16(x, ffn, p_drop):
    s = ffn.sequential
    return _FfnHalfBf16.apply(x.contiguous(), s[0].weight, s[0].bias, s[0].eps, s[1].weight, s[1].bias, s[4].weight, s[4].bias, float(p_drop))


class _PostNetBf16(torch.autograd.Function):
    """PostNet (fs2/layers.py:143-212: 5 × [Conv1d k5 → BatchNorm1d → tanh (not on the last) → Dropout 0.5]) in the bf16 mode
    as ONE autograd node.  Between the layers the activations exist only in bf16 (written by the BatchNorm-affine + tanh
    + dropout kernel, read by the next convolution's TMA), and in the backward the BatchNorm/tanh gradient gz exists only in
    bf16 (read by TMA in both the data-gradient and the weight-gradient contraction): the five k = 5, 512-channel
    convolutions — the compute-bound third of the model — run TMA-fed in all three directions."""

    @staticmethod
    def forward(ctx, x, training, p_drop, bns, *params):
        n = len(bns)
        h = x
        saved, meta = [], []
        for i in range(n):
            weight, bias = params[4 * i], params[4 * i + 1]
            w_taps = ops.conv_weight_taps(weight)
            taps, N, K = w_taps.shape
            pad = (taps - 1) // 2
            w16, _ = ops.bf16_weight(w_taps)
            hint = 256 if (taps * K >= 1024 and N % 256 == 0) else 0
            z, _, _ = ops.gemm_bf16(h, w16, bias.detach() if bias is not None else None, taps_pad=pad, block_n_hint=hint)
            scale, shift, mean, rstd = ops.bn_scale_shift(bns[i], z, training, save_stats=True)
            seed = _seed(p_drop)
            act = "tanh" if i < n - 1 else None
            saved += [h, w16, z, scale, shift, mean, rstd]
            meta.append((act, pad, seed, hint, bias is not None))
            if i < n - 1:
                h = ops.affine_act(z, scale, shift, act, None, p_drop, seed, out_bf16=True)
            else:
                out = ops.affine_act(z, scale, shift, act, None, p_drop, seed)
        ctx.save_for_backward(*saved)
        ctx.meta, ctx.cfg, ctx.params = meta, (training, p_drop, n), params
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        training, p_drop, n = ctx.cfg
        saved, params = ctx.saved_tensors, ctx.params
        grads = [None] * (4 * n)
        g = g.contiguous()
        sink = _SINK
        for i in range(n - 1, -1, -1):
            h, w16, z, scale, shift, mean, rstd = saved[7 * i: 7 * i + 7]
            act, pad, seed, hint, has_bias = ctx.meta[i]
            weight, bias, bn_w, bn_b = params[4 * i: 4 * i + 4]
            taps = w16.shape[0]
            direct_bn = _direct_grads(bn_w, bn_b)
            gz16, dgamma, dbeta = ops.bn_act_bwd(g, z, scale, shift, mean, rstd, act, training, p_drop, seed,
                                                 accumulate_into=direct_bn, out_bf16=True)
            if direct_bn is None:
                grads[4 * i + 2], grads[4 * i + 3] = dgamma, dbeta
            direct_w = _direct_grads(weight, bias) if has_bias else _direct_grads(weight)
            if sink is not None and direct_w is not None:
                sink.stream.wait_stream(torch.cuda.current_stream())
                sink.keep.append((gz16, h))
                with torch.cuda.stream(sink.stream):
                    if has_bias:
                        ops.colsum(gz16, out=bias.grad, accumulate=True)
                    ops.gemm_wgrad_bf16(gz16, h, taps, pad, True, accumulate_into=weight.grad)
            else:
                if has_bias:
                    grads[4 * i + 1] = ops.colsum(gz16)
                grads[4 * i] = ops.gemm_wgrad_bf16(gz16, h, taps, pad, True)
            if i > 0 or ctx.needs_input_grad[0]:
                g, _, _ = ops.gemm_bf16(gz16, w16, None, w_mn=True, taps_pad=taps - 1 - pad, block_n_hint=hint)
        return (g if ctx.needs_input_grad[0] else None, None, None, None, *grads)


def postnet_bf16(x, pn, training, dropout_p):
    params = []
    bns = []
    for block in pn.convolutions:
        conv, bn = block[0].conv, block[1]
        params += [conv.weight, conv.bias, bn.weight, bn.bias]
        bns.append(bn)
    return _PostNetBf16.apply(x.contiguous(), bool(training), float(dropout_p), bns, *params)


class _GluDwconvBnSilu(torch.autograd.Function):
    """GLU → depthwise conv → BatchNorm1d → SiLU (torchaudio conformer.py:50-65)."""

    @staticmethod
    def forward(ctx, h, dw_w, dw_b, bn_w, bn_b, bn, training):
        C = dw_w.shape[0]
        z = ops.dwconv(h, dw_w, dw_b, channels=C, glu=True)
        scale, shift, mean, rstd = ops.bn_scale_shift(bn, z, training, save_stats=True)
        y = ops.affine_act(z, scale, shift, "silu")
        ctx.save_for_backward(h, dw_w, z, scale, shift, mean, rstd)
        ctx.training = training
        ctx.params = (dw_w, dw_b, bn_w, bn_b)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        h, dw_w, z, scale, shift, mean, rstd = ctx.saved_tensors
        dw_p, db_p, bnw_p, bnb_p = ctx.params
        direct_bn, direct_dw = _direct_grads(bnw_p, bnb_p), _direct_grads(dw_p, db_p)
        gz, dgamma, dbeta = ops.bn_act_bwd(g.contiguous(), z, scale, shift, mean, rstd, "silu", ctx.training, accumulate_into=direct_bn)
        dh, ddw, ddb = ops.dwconv_bwd(gz, h, dw_w, glu=True, accumulate_into=direct_dw)
        if direct_bn is not None:
            dgamma = dbeta = None
        if direct_dw is not None:
            ddw = ddb = None
        return dh, ddw, ddb, dgamma, dbeta, None, None


def glu_dwconv_bn_silu(h, dw_w, dw_b, bn, training):
    return _GluDwconvBnSilu.apply(h.contiguous(), dw_w, dw_b, bn.weight, bn.bias, bn, training)


class _Dwconv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.params = (weight, bias)
        return ops.dwconv(x, weight, bias, channels=weight.shape[0])

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        direct = _direct_grads(*ctx.params)
        dx, dw, db = ops.dwconv_bwd(g.contiguous(), x, weight, glu=False, accumulate_into=direct)
        return (dx, None, None) if direct is not None else (dx, dw, db)


def dwconv(x, weight, bias):
    return _Dwconv.apply(x.contiguous(), weight, bias)


# ---------------------------------------------------------------------------------------------
class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, lengths, heads, dropout_p, order):
        seed = int(torch.randint(0, 2**31 - 1, (1,)).item()) if dropout_p else 0
        out, lse = ops.attention(qkv, lengths, heads, want_lse=True, dropout_p=dropout_p, seed=seed, order=order)
        ctx.save_for_backward(qkv, out, lse, lengths)
        ctx.meta = (heads, dropout_p, seed)
        ctx.order = order
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        qkv, out, lse, lengths = ctx.saved_tensors
        heads, p, seed = ctx.meta
        return ops.attention_bwd(qkv, out, lse, g.contiguous(), lengths, heads, p, seed, ctx.order), None, None, None, None


def attention(qkv, lengths, heads, dropout_p=0.0, order=None):
    return _Attention.apply(qkv.contiguous(), lengths, heads, float(dropout_p), order)


class _QkvAttention(torch.autograd.Function):
    """bf16 mode: in-projection GEMM (bf16 qkv straight from its epilogue) + tensor-core attention as one node —
    the fp32 qkv tensor never exists (torchaudio conformer.py:193-202 / nn.MultiheadAttention's in_proj + SDPA)."""

    @staticmethod
    def forward(ctx, x, weight, bias, lengths, heads, dropout_p, order):
        seed = _seed(dropout_p)
        w16, _ = ops.bf16_weight(weight.detach())
        _, qkv16, _ = ops.gemm_bf16(x, w16, bias.detach() if bias is not None else None, want_c=False, want_c16=True)
        out, lse = ops.attention_bf16(qkv16, lengths, heads, want_lse=True, dropout_p=dropout_p, seed=seed, order=order)
        ctx.save_for_backward(x, qkv16, out, lse, lengths)
        ctx.meta = (heads, dropout_p, seed, bias is not None)
        ctx.order = order
        ctx.params = (weight, bias)
        ctx.precision = ops.PRECISION
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, qkv16, out, lse, lengths = ctx.saved_tensors
        heads, p, seed, has_bias = ctx.meta
        weight, bias = ctx.params
        dqkv = ops.attention_bwd_bf16(qkv16, out, lse, g.contiguous(), lengths, heads, p, seed, ctx.order)
        w_taps = weight.detach().reshape(1, *weight.shape)
        dx, dw, db = _gemm_backward(dqkv, x, w_taps, 0, False, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                    has_bias and ctx.needs_input_grad[2], weight, bias, wt=None, precision=ctx.precision)
        return dx, dw, db, None, None, None, None


def qkv_attention_bf16(x, weight, bias, lengths, heads, dropout_p=0.0, order=None):
    return _QkvAttention.apply(x.contiguous(), weight, bias, lengths, heads, float(dropout_p), order)


class _Rowdot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, mask):
        ctx.save_for_backward(x, weight, mask)
        return ops.rowdot(x, weight.reshape(-1), bias, mask)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, weight, mask = ctx.saved_tensors
        dx, dw, db = ops.rowdot_bwd(g.contiguous(), x, weight.reshape(-1), mask)
        return dx, dw.reshape(weight.shape), db, None


def rowdot(x, weight, bias, mask):
    return _Rowdot.apply(x.contiguous(), weight, bias, mask)


class _AlignerScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, prior, key_lens):
        soft, logprob = ops.aligner_scores(q, k, prior, key_lens)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(q, k, prior, key_lens, soft, logprob)
        return soft, logprob

    @staticmethod
    @once_differentiable
    def backward(ctx, g_soft, g_logprob):
        q, k, prior, key_lens, soft, logprob = ctx.saved_tensors
        if g_soft is None and g_logprob is None:
            return None, None, None, None
        dq, dk = ops.aligner_bwd(g_soft, g_logprob, soft, logprob, prior, key_lens, q, k)
        return dq, dk, None, None


def aligner_scores(q, k, prior, key_lens):
    return _AlignerScores.apply(q.contiguous(), k.contiguous(), prior, key_lens)


# ---------------------------------------------------------------------------------------------
class _LengthRegulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cum, total, width, inv_freq):
        out, out_pos, mask, _ = ops.lr_gather(x, cum, total, width, inv_freq, want_out=True)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(cum)
        ctx.meta = (x.shape[1], x.shape[2], width, out_pos is not None)
        ctx.mark_non_differentiable(mask)
        if out_pos is None:
            return out, mask
        return out, out_pos, mask

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        (cum,) = ctx.saved_tensors
        T, D, width, has_pos = ctx.meta
        g_out = grads[0]
        g_pos = grads[1] if has_pos else None
        if g_out is None and g_pos is None:
            return None, None, None, None, None
        dx = ops.lr_bwd(g_out, g_pos, cum, T, D, width)
        return dx, None, None, None, None


def length_regulate(x, cum, total, width, inv_freq):
    if torch.is_grad_enabled() and x.requires_grad:
        r = _LengthRegulate.apply(x.contiguous(), cum, total, width, inv_freq)
        return (r[0], None, r[1]) if inv_freq is None else r
    out, out_pos, mask, _ = ops.lr_gather(x, cum, total, width, inv_freq, want_out=True)
    return out, out_pos, mask


class _BucketizeEmbedAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, table, v, scale, bins):
        y, ids, _ = ops.bucketize_embed_add(v, bins, table, x, scale=scale, want_ids=True)
        ctx.save_for_backward(ids)
        ctx.table_shape = table.shape
        ctx.mark_non_differentiable(ids)
        return y, ids

    @staticmethod
    @once_differentiable
    def backward(ctx, g, _gids):
        (ids,) = ctx.saved_tensors
        g = g.contiguous()
        dtable = None
        if ctx.needs_input_grad[1]:
            dtable = torch.zeros(ctx.table_shape, dtype=torch.float32, device=g.device)
            ops.scatter_add_rows(dtable, g, ids)
        return g, dtable, None, None, None


def bucketize_embed_add(v, scale, bins, table, x, return_scaled=False):
    if return_scaled:  # inference: the scaled prediction is returned; no gradient flows here
        y, ids, vs = ops.bucketize_embed_add(v.detach().contiguous(), bins, _d(table), _d(x), scale=scale, want_ids=True, want_scaled=True)
        return y, ids, vs
    if torch.is_grad_enabled() and (x.requires_grad or table.requires_grad):
        return _BucketizeEmbedAdd.apply(x.contiguous(), table, v.detach().contiguous(), float(scale), bins.detach())
    y, ids, _ = ops.bucketize_embed_add(v.detach().contiguous(), bins, table, x, scale=scale, want_ids=True)
    return y, ids


def embedding_lookup(ids, table):
    return _EmbeddingLookup.apply(table, ids)


class _EmbeddingLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, ids):
        ctx.save_for_backward(ids)
        ctx.table_shape = table.shape
        return ops.gather_rows(table, ids)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        dtable = torch.zeros(ctx.table_shape, dtype=torch.float32, device=g.device)
        ops.scatter_add_rows(dtable, g.contiguous(), ids)
        return dtable, None


class _Scale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, factor):
        ctx.factor = factor
        return ops.axpby(x, factor)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return ops.axpby(g.contiguous(), ctx.factor), None


def scale(x, factor):
    return _Scale.apply(x.contiguous(), float(factor))


class _EmbedPosenc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, text, inv_freq, lens, padding_idx):
        emb, x = ops.embed_posenc(text, table, inv_freq, lens)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(text)
        ctx.meta = (table.shape, -1 if padding_idx is None else int(padding_idx))
        return emb, x

    @staticmethod
    @once_differentiable
    def backward(ctx, g_emb, g_x):
        (text,) = ctx.saved_tensors
        shape, pad = ctx.meta
        ref = g_emb if g_emb is not None else g_x
        if ref is None:
            return None, None, None, None, None
        dtable = torch.zeros(shape, dtype=torch.float32, device=ref.device)
        g1 = g_x if g_x is not None else g_emb
        g2 = g_emb if (g_x is not None and g_emb is not None) else None
        ops.scatter_add_rows(dtable, g1.contiguous(), text, None if g2 is None else g2.contiguous(), skip_id=pad)
        return dtable, None, None, None, None


def embed_posenc(text, table, inv_freq, lens, padding_idx):
    text = text.contiguous()
    if text.dtype != torch.int32:
        text = text.to(torch.int32)
    if torch.is_grad_enabled() and table.requires_grad:
        return _EmbedPosenc.apply(table, text, inv_freq, lens, padding_idx)
    return ops.embed_posenc(text, table, inv_freq, lens)


class _AddPosenc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inv_freq, lens):
        return ops.add_posenc(x, inv_freq, lens)

    @staticmethod
    def backward(ctx, g):
        return g, None, None


def add_posenc(x, inv_freq, lens):
    return _AddPosenc.apply(x.contiguous(), inv_freq, lens)


class _AddRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_rows, *rows_and_ids):
        rows = rows_and_ids[:n_rows]
        ids = rows_and_ids[n_rows:]
        ctx.save_for_backward(*[i for i in ids if i is not None])
        ctx.meta = (n_rows, [r.shape for r in rows], [i is not None for i in ids])
        return ops.add_rows(x, list(zip([r.contiguous() for r in rows], ids)))

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        n_rows, shapes, has_ids = ctx.meta
        saved = list(ctx.saved_tensors)
        g = g.contiguous()
        grads = []
        for k in range(n_rows):
            ids = saved.pop(0) if has_ids[k] else None
            if ctx.needs_input_grad[2 + k]:
                d = torch.zeros(shapes[k], dtype=torch.float32, device=g.device)
                ops.rows_sum_scatter(d, g, ids)
                grads.append(d)
            else:
                grads.append(None)
        return (g, None, *grads, *([None] * n_rows))


def add_rows(x, rows):
    return _AddRows.apply(x.contiguous(), len(rows), *[r for r, _ in rows], *[i for _, i in rows])


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.axpby(a, 1.0, b, 1.0)

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a, b):
    return _Add.apply(a.contiguous(), b.contiguous())


# ---------------------------------------------------------------------------------------------
class _MaskedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, kind, weight, log1p_int_target):
        ctx.save_for_backward(pred, target, mask)
        ctx.meta = (kind, weight, log1p_int_target)
        return ops.masked_loss_fwd(pred, target, mask, kind, weight, log1p_int_target)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        pred, target, mask = ctx.saved_tensors
        kind, weight, log1p = ctx.meta
        return ops.masked_loss_bwd(pred, target, mask, kind, weight, g.contiguous(), log1p), None, None, None, None, None


def masked_loss(pred, target, mask, kind, weight, log1p_int_target=False):
    return _MaskedLoss.apply(pred.contiguous(), target.detach().contiguous(), mask, kind, float(weight), bool(log1p_int_target))


class _BinLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hard, soft, eps):
        loss, sums = ops.bin_loss_fwd(hard, soft, eps)
        ctx.save_for_backward(hard, soft, sums)
        ctx.eps = eps
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        hard, soft, sums = ctx.saved_tensors
        return None, ops.bin_loss_bwd(hard, soft, ctx.eps, sums, g.contiguous()), None


def attn_bin_loss(hard, soft, eps=1e-12):
    return _BinLoss.apply(hard.detach().contiguous(), soft.contiguous(), float(eps))


class _CtcForwardSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attn_logprob, key_lens, query_lens, blank_logprob):
        loss, saved = ops.ctc_forward_sum_fwd(attn_logprob, key_lens, query_lens, blank_logprob)
        ctx.save_for_backward(*saved)
        ctx.blank = blank_logprob
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return ops.ctc_forward_sum_bwd(ctx.saved_tensors, g.contiguous(), ctx.blank), None, None, None


def ctc_forward_sum(attn_logprob, key_lens, query_lens, blank_logprob=-1.0):
    pre = getattr(attn_logprob, "_fs2k_ctc", None)
    if pre is not None and pre[2] == float(blank_logprob) and pre[3] is key_lens and pre[4] is query_lens:
        torch.cuda.current_stream().wait_stream(pre[1])  # started early on the side stream (ctc_forward_sum_prefetch)
        return pre[0]
    return _CtcForwardSum.apply(attn_logprob.contiguous(), key_lens, query_lens, float(blank_logprob))


def ctc_forward_sum_prefetch(attn_logprob, key_lens, query_lens, blank_logprob=-1.0):
    """With a WgradSink installed: start the forward-sum loss on the side stream as soon as the aligner has produced
    attn_logprob.  Its sequential recursion (one CTA per utterance, F dependent steps) then overlaps the encoder /
    decoder forward, and — autograd runs a node's backward on its forward's stream — its backward overlaps the
    decoder backward.  The loss module picks the result up from the tensor (`ctc_forward_sum`)."""
    sink = _SINK
    if sink is None or not torch.is_grad_enabled() or not attn_logprob.requires_grad:
        return
    sink.stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(sink.stream):
        loss = _CtcForwardSum.apply(attn_logprob.contiguous(), key_lens, query_lens, float(blank_logprob))
    attn_logprob._fs2k_ctc = (loss, sink.stream, float(blank_logprob), key_lens, query_lens)


# ---------------------------------------------------------------------------------------------
# GST reference encoder, training path (fs2/gst/model.py:103-257)
# ---------------------------------------------------------------------------------------------
class _Conv2dS2BnRelu(torch.autograd.Function):
    """Conv2d(3×3, stride 2, pad 1, no bias) → BatchNorm2d → ReLU on channels-last x [B,H,W,Ci]."""

    @staticmethod
    def forward(ctx, x, weight, bn_w, bn_b, bn, training):
        w = weight.detach().permute(2, 3, 1, 0).contiguous()  # [Co,Ci,3,3] → [3,3,Ci,Co]
        z = ops.conv2d_s2_raw(x, w)
        z2 = z.view(-1, z.shape[-1])
        scale, shift, mean, rstd = ops.bn_scale_shift(bn, z2, training, save_stats=True)
        y = ops.affine_act(z2, scale, shift, "relu").view_as(z)
        ctx.save_for_backward(x, w, z, scale, shift, mean, rstd)
        ctx.training = training
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w, z, scale, shift, mean, rstd = ctx.saved_tensors
        Co = z.shape[-1]
        gz, dgamma, dbeta = ops.bn_act_bwd(g.contiguous().view(-1, Co), z.view(-1, Co), scale, shift, mean, rstd, "relu", ctx.training)
        gz = gz.view_as(z)
        dx = ops.conv2d_s2_dgrad(gz, w, x.shape) if ctx.needs_input_grad[0] else None
        dw = ops.conv2d_s2_wgrad(x, gz).permute(3, 2, 0, 1).contiguous()
        return dx, dw, dgamma, dbeta, None, None


def conv2d_s2_bn_relu(x, conv, bn, training):
    return _Conv2dS2BnRelu.apply(x.contiguous(), conv.weight, bn.weight, bn.bias, bn, training)


class _GruLastHidden(torch.autograd.Function):
    """Last hidden state of a one-layer batch_first GRU (h0 = 0) with back-propagation through time."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh):
        B, T, I = x.shape
        U = w_hh.shape[1]
        x2 = x.reshape(B * T, I)
        xp = ops.gemm(x2, w_ih.detach(), b_ih.detach())
        h = torch.zeros((B, U), dtype=torch.float32, device=x.device)
        hs, hps = [], []
        for t in range(T):
            hp = ops.gemm(h, w_hh.detach(), b_hh.detach())
            hs.append(h)
            hps.append(hp)
            h = ops.gru_gate(xp, t, T, hp, h)
        ctx.save_for_backward(x2, w_ih, w_hh, xp, torch.stack(hs), torch.stack(hps))
        ctx.shape = (B, T, I, U)
        ctx.precision = ops.PRECISION
        return h

    @staticmethod
    @once_differentiable
    def backward(ctx, dh):
        with ops.precision_scope(ctx.precision):
            return _GruLastHidden._backward(ctx, dh)

    @staticmethod
    def _backward(ctx, dh):
        x2, w_ih, w_hh, xp, hs, hps = ctx.saved_tensors
        B, T, I, U = ctx.shape
        dxp = torch.empty_like(xp)
        dw_hh = torch.zeros_like(w_hh)
        db_hh = torch.zeros((3 * U,), dtype=torch.float32, device=dh.device)
        wt_hh = ops.weight_taps_transposed(w_hh.detach().reshape(1, 3 * U, U))  # [1, U, 3U]
        dh = dh.contiguous()
        for t in range(T - 1, -1, -1):
            dhp, dh_prev = ops.gru_gate_bwd(xp, dxp, t, T, hps[t], hs[t], dh)
            ops.gemm_wgrad(dhp, hs[t], 1, 0, False, accumulate_into=dw_hh)
            ops.colsum(dhp, out=db_hh, accumulate=True)
            dh = ops.gemm(dhp, wt_hh, None, residual=dh_prev)
        dw_ih = ops.gemm_wgrad(dxp, x2, 1, 0, False)
        db_ih = ops.colsum(dxp)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dxp, ops.weight_taps_transposed(w_ih.detach().reshape(1, 3 * U, I)), None).view(B, T, I)
        return dx, dw_ih, dw_hh, db_ih, db_hh


def gru_last_hidden(x, gru):
    return _GruLastHidden.apply(x.contiguous(), gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0)


class _GstTokenAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, heads):
        ctx.save_for_backward(q, k, v)
        ctx.heads = heads
        return ops.gst_token_attention(q, k, v, heads)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        q, k, v = ctx.saved_tensors
        dq, dk, dv = ops.gst_token_attention_bwd(q, k, v, g.contiguous(), ctx.heads)
        return dq, dk, dv, None


def gst_token_attention(q, k, v, heads):
    return _GstTokenAttention.apply(q.contiguous(), k.contiguous(), v.contiguous(), heads)


class _Tanh(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = ops.tanh(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return ops.act_bwd(g.contiguous(), y, "tanh")


def tanh(x):
    if torch.is_grad_enabled() and x.requires_grad:
        return _Tanh.apply(x.contiguous())
    return ops.tanh(x.detach())


def tanh_row(table, index):
    return ops.tanh(table.detach()[index: index + 1].contiguous())
