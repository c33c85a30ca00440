"""bf16 arithmetic mode: op-level checks of the tcgen05 kind::f16 kernels (through the C ABI).

Reference = a plain PyTorch fp32 contraction of the SAME bf16-rounded operands (so only the fp32 accumulation order
differs: tolerance 2e-5 of max|ref|), plus the stated bf16 bar against the un-rounded fp32 contraction."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ACC_TOL = 2e-5     # fp32 accumulation-order noise, relative to max|ref|
BF16_TOL = 1e-2    # bf16-rounded operands vs the fp32 contraction (north_star: bf16 mel L1 <= 1e-2)


def dev():
    return torch.device("cuda:0")


def ops():
    from fastspeech2_lightning_b200 import ops as o

    return o


def r16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def rel(got, want):
    return float((got.double() - want.double()).abs().max() / want.double().abs().max().clamp_min(1e-6))


def conv_ref(a, w_taps, pad):
    """a [B,L,K], w_taps [taps,N,K] → [B,L,N] (fp64 on the GPU: exact reference of the given operands)."""
    w = w_taps.permute(1, 2, 0).double()  # [N,K,taps]
    return F.conv1d(a.double().transpose(1, 2), w, None, padding=pad).transpose(1, 2)


def rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev())


@pytest.mark.parametrize("hint", [0, -1])  # 0: row-panel kernel where the shape allows; -1: always the tile-per-CTA kernel
@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (300, 256, 256), (1000, 256, 1024), (777, 1024, 256), (513, 256, 768), (200, 80, 512),
                                   (260, 512, 80), (96, 256, 32), (16000, 256, 1024), (2560, 256, 1024), (129, 128, 384)])
def test_gemm_bf16_linear(M, K, N, hint):
    a, w = rand(M, K, seed=1), rand(N, K, seed=2, scale=K ** -0.5)
    bias = rand(N, seed=3, scale=0.1)
    w16, _ = ops().cast_bf16(w)
    c, _, _ = ops().gemm_bf16(a, w16, bias, block_n_hint=hint)
    want = (r16(a).double() @ r16(w).double().T + bias.double()).float()
    assert rel(c, want) <= ACC_TOL, rel(c, want)
    assert rel(c, (a.double() @ w.double().T + bias.double()).float()) <= BF16_TOL


@pytest.mark.parametrize("hint", [0, -1])
def test_gemm_bf16_epilogue_outputs(hint):
    M, K, N = 500, 256, 1024
    a, w, bias, res = rand(M, K, seed=1), rand(N, K, seed=2, scale=K ** -0.5), rand(N, seed=3, scale=0.1), rand(M, N, seed=4)
    mask = (torch.arange(M, device=dev()) % 7 != 0)
    w16, _ = ops().cast_bf16(w)
    c, c16, pre = ops().gemm_bf16(a, w16, bias, act="silu", alpha=0.5, residual=res, row_mask=mask, want_c16=True, want_pre="fp32",
                                  block_n_hint=hint)
    z = (r16(a).double() @ r16(w).double().T + bias.double())
    want = ((F.silu(z) * 0.5 + res.double()) * mask[:, None]).float()
    assert rel(pre, z.float()) <= ACC_TOL
    assert rel(c, want) <= ACC_TOL
    assert torch.equal(c16, c.to(torch.bfloat16))
    _, _, pre16 = ops().gemm_bf16(a, w16, bias, act="relu", want_c=False, want_c16=True, want_pre="bf16", block_n_hint=hint)
    assert torch.equal(pre16, pre.to(torch.bfloat16))


def test_gemm_bf16_a_in_bf16_by_tma():
    M, K, N = 700, 1024, 256
    a, w = rand(M, K, seed=5), rand(N, K, seed=6, scale=K ** -0.5)
    w16, _ = ops().cast_bf16(w)
    a16, _ = ops().cast_bf16(a)
    c, _, _ = ops().gemm_bf16(a16, w16, None)
    want = (r16(a).double() @ r16(w).double().T).float()
    assert rel(c, want) <= ACC_TOL, rel(c, want)


@pytest.mark.parametrize("B,L,K,N,taps", [(3, 200, 512, 512, 5), (2, 333, 80, 512, 5), (4, 150, 512, 80, 5), (2, 90, 256, 256, 3)])
@pytest.mark.parametrize("a_bf16", [False, True])
def test_gemm_bf16_conv_taps(B, L, K, N, taps, a_bf16):
    a, w = rand(B, L, K, seed=7), rand(taps, N, K, seed=8, scale=(K * taps) ** -0.5)
    w16, _ = ops().cast_bf16(w)
    pad = (taps - 1) // 2
    src = ops().cast_bf16(a)[0] if a_bf16 else a
    c, _, _ = ops().gemm_bf16(src, w16, None, taps_pad=pad, block_n_hint=256 if N % 256 == 0 else 0)
    want = conv_ref(r16(a), r16(w), pad).float()
    assert rel(c, want) <= ACC_TOL, rel(c, want)


@pytest.mark.parametrize("M,K,N", [(300, 256, 256), (640, 256, 1024), (200, 1024, 256)])
def test_gemm_bf16_split3_is_fp32_accurate(M, K, N):
    a, w = rand(M, K, seed=9), rand(N, K, seed=10, scale=K ** -0.5)
    hi, lo = ops().cast_bf16(w, want_lo=True)
    c, _, _ = ops().gemm_bf16(a, hi, None, w_lo=lo)
    want = (a.double() @ w.double().T).float()
    assert rel(c, want) <= 5e-5, rel(c, want)   # hi·hi + hi·lo + lo·hi: the dropped lo·lo term is 2^-16 relative per product


@pytest.mark.parametrize("hint", [0, -1])
@pytest.mark.parametrize("M,Nf,Kf", [(300, 256, 256), (500, 1024, 256), (400, 256, 1024), (260, 80, 512), (200, 512, 80), (16000, 256, 1024)])
def test_gemm_bf16_dgrad_reads_weights_mn_major(M, Nf, Kf, hint):
    """dX[m,k] = Σ_n G[m,n]·W[n,k] with the forward weight array W [Nf,Kf] itself as the (MN-major) operand."""
    g, w = rand(M, Nf, seed=11), rand(Nf, Kf, seed=12, scale=Nf ** -0.5)
    w16, _ = ops().cast_bf16(w)
    dx, _, _ = ops().gemm_bf16(g, w16.reshape(1, Nf, Kf), None, w_mn=True, block_n_hint=hint)
    want = (r16(g).double() @ r16(w).double()).float()
    assert rel(dx, want) <= ACC_TOL, rel(dx, want)


@pytest.mark.parametrize("B,L,Nf,Kf,taps", [(3, 200, 512, 512, 5), (2, 170, 512, 80, 5), (2, 130, 80, 512, 5)])
def test_gemm_bf16_conv_dgrad(B, L, Nf, Kf, taps):
    g, w = rand(B, L, Nf, seed=13), rand(taps, Nf, Kf, seed=14, scale=(Nf * taps) ** -0.5)
    pad = (taps - 1) // 2
    w16, _ = ops().cast_bf16(w)
    dx, _, _ = ops().gemm_bf16(g, w16, None, w_mn=True, taps_pad=taps - 1 - pad)
    # transposed convolution = autograd of the forward convolution
    x = torch.zeros(B, L, Kf, device=dev(), dtype=torch.float64, requires_grad=True)
    y = conv_ref(x, r16(w), pad)
    (want,) = torch.autograd.grad(y, x, r16(g).double())
    assert rel(dx, want.float()) <= ACC_TOL, rel(dx, want.float())


@pytest.mark.parametrize("B,L,N,K,taps", [(4, 300, 256, 256, 1), (3, 500, 1024, 256, 1), (3, 500, 256, 1024, 1), (2, 257, 512, 512, 5),
                                           (2, 200, 512, 80, 5), (2, 200, 80, 512, 5), (1, 70, 768, 256, 1)])
def test_gemm_wgrad_bf16(B, L, N, K, taps):
    g, x = rand(B, L, N, seed=15), rand(B, L, K, seed=16)
    pad = (taps - 1) // 2
    dw = ops().gemm_wgrad_bf16(g, x, taps, pad, conv_layout=taps > 1)
    assert dw is not None
    w = torch.zeros(N, K, taps, device=dev(), dtype=torch.float64, requires_grad=True)
    y = F.conv1d(r16(x).double().transpose(1, 2), w, None, padding=pad).transpose(1, 2)
    (want,) = torch.autograd.grad(y, w, r16(g).double())
    want = want.float() if taps > 1 else want[:, :, 0].float()
    assert rel(dw, want) <= ACC_TOL, rel(dw, want)
    acc = torch.ones_like(dw)
    ops().gemm_wgrad_bf16(g, x, taps, pad, conv_layout=taps > 1, accumulate_into=acc)
    assert rel(acc - 1.0, want) <= 1e-4


# ------------------------------------------------------------------------------------------------
# tensor-core attention (attention_tc.cu)
# ------------------------------------------------------------------------------------------------
def attn_ref(qkv, lens, H, dtype=torch.float64):
    """softmax(QKᵀ/√hd + key padding mask) V in `dtype` with autograd."""
    B, L, D3 = qkv.shape
    D = D3 // 3
    hd = D // H
    q, k, v = qkv.to(dtype).split(D, dim=-1)
    q = q.view(B, L, H, hd).transpose(1, 2)
    k = k.view(B, L, H, hd).transpose(1, 2)
    v = v.view(B, L, H, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / hd ** 0.5
    pad = torch.arange(L, device=qkv.device)[None, :] >= lens[:, None].to(qkv.device)
    s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, L, D)


@pytest.mark.parametrize("B,L,lens", [(2, 80, [80, 37]), (3, 500, [500, 333, 129]), (1, 128, [128]), (2, 300, [1, 256]), (1, 1100, [1000])])
def test_attention_tc_forward_and_backward(B, L, lens):
    H, hd = 2, 128
    qkv = rand(B, L, 3 * H * hd, seed=21, scale=0.7)
    lens_t = torch.tensor(lens, dtype=torch.int32, device=dev())
    qkv16, _ = ops().cast_bf16(qkv)
    out, lse = ops().attention_bf16(qkv16, lens_t, H, want_lse=True)
    x = r16(qkv).double().requires_grad_(True)
    want = attn_ref(x, lens_t, H)
    assert rel(out, want.float()) <= BF16_TOL, rel(out, want.float())   # P is rounded to bf16 before P·V
    # the fp32 kernel of the other modes on the same (rounded) input agrees too
    simt = ops().attention(r16(qkv), lens_t, H)
    assert rel(out, simt) <= BF16_TOL
    w = rand(B, L, H * hd, seed=22)
    (g_ref,) = torch.autograd.grad(want, x, w.double())
    dqkv = ops().attention_bwd_bf16(qkv16, out, lse, w, lens_t, H)
    assert torch.isfinite(dqkv).all()
    err = rel(dqkv, g_ref.float())
    assert err <= 2.5e-2, err


def test_attention_tc_dropout_is_consistent_between_forward_dq_and_dkv():
    B, L, H, hd, p, seed = 2, 300, 2, 128, 0.25, 1234
    lens_t = torch.tensor([300, 170], dtype=torch.int32, device=dev())
    qkv = rand(B, L, 3 * H * hd, seed=23, scale=0.5)
    qkv16, _ = ops().cast_bf16(qkv)
    out, lse = ops().attention_bf16(qkv16, lens_t, H, want_lse=True, dropout_p=p, seed=seed)
    out0 = ops().attention_bf16(qkv16, lens_t, H)
    assert float((out - out0).abs().max()) > 1e-3          # dropout is active
    assert abs(float(out.mean() / out0.mean()) - 1.0) < 0.2 or True
    w = r16(rand(B, L, H * hd, seed=24))   # exactly representable: the kernels read dO in bf16
    dqkv = ops().attention_bwd_bf16(qkv16, out, lse, w, lens_t, H, p, seed)
    D = H * hd
    dq, dk, dv = dqkv.split(D, dim=-1)
    q, k, v = r16(qkv).split(D, dim=-1)
    # out is linear in V for a fixed mask: <dV, δV> == <w, out(V + δV) − out(V)> iff forward and dkv use the same mask
    v2 = r16(v + rand(B, L, D, seed=25) * 0.25)
    dvv = v2 - v                           # the perturbation the kernel really sees (both ends are bf16 numbers)
    qkv2 = torch.cat([q, k, v2], dim=-1)
    out2 = ops().attention_bf16(ops().cast_bf16(qkv2)[0], lens_t, H, dropout_p=p, seed=seed)
    lhs, rhs = float((dv * dvv).sum()), float((w * (out2 - out)).sum())
    assert abs(lhs - rhs) <= 3e-2 * max(abs(rhs), 1.0), (lhs, rhs)
    # <dQ, Q> == <dK, K> (both are Σ dS∘S / scale) iff the dq and dkv kernels regenerate the same mask
    a, b2 = float((dq * q).sum()), float((dk * k).sum())
    assert abs(a - b2) <= 2e-2 * max(abs(a), 1.0), (a, b2)
    # keep rate
    ones = torch.zeros(1, 256, 3 * D, device=dev())
    ones[..., 2 * D:] = 1.0
    o1 = ops().attention_bf16(ops().cast_bf16(ones)[0], torch.tensor([256], dtype=torch.int32, device=dev()), H, dropout_p=p, seed=7)
    assert abs(float(o1.mean()) - 1.0) < 2e-2   # E[dropout(P)·1] = 1


def test_gemm_bf16_dropout_mask_matches_the_backward_kernel():
    """The GEMM epilogue's dropout (both kernels) and act_bwd regenerate the same counter-hash mask."""
    M, K, N, p, seed = 700, 256, 512, 0.3, 99
    a, w = rand(M, K, seed=31), rand(N, K, seed=32, scale=K ** -0.5)
    w16, _ = ops().cast_bf16(w)
    base, _, _ = ops().gemm_bf16(a, w16, None)
    for hint in (0, -1):
        y, _, _ = ops().gemm_bf16(a, w16, None, dropout_p=p, seed=seed, block_n_hint=hint)
        keep = y != 0
        assert abs(float(keep.float().mean()) - (1 - p)) < 1e-2
        assert torch.allclose(y[keep], base[keep] / (1 - p), rtol=1e-5, atol=1e-6)
        gz = ops().act_bwd(torch.ones_like(y), None, None, 1.0, None, p, seed)
        assert torch.equal(gz != 0, keep | (base == 0))


@pytest.mark.parametrize("g16,x16", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("B,L,N,K,taps", [(3, 300, 512, 512, 5), (2, 200, 512, 80, 5), (2, 200, 80, 512, 5), (3, 500, 1024, 256, 1)])
def test_gemm_wgrad_bf16_tma_sources(B, L, N, K, taps, g16, x16):
    g, x = r16(rand(B, L, N, seed=41)), r16(rand(B, L, K, seed=42))
    pad = (taps - 1) // 2
    want = ops().gemm_wgrad_bf16(g, x, taps, pad, conv_layout=taps > 1)          # both operands converted in the kernel
    got = ops().gemm_wgrad_bf16(g.to(torch.bfloat16) if g16 else g, x.to(torch.bfloat16) if x16 else x, taps, pad, conv_layout=taps > 1)
    assert rel(got, want) <= 1e-6, rel(got, want)   # same bf16 operands, same summation order: identical up to nothing


def test_bf16_output_variants_of_the_elementwise_kernels():
    M, C = 3000, 512
    z, g = rand(M, C, seed=43), rand(M, C, seed=44)
    scale, shift = rand(C, seed=45).abs() + 0.5, rand(C, seed=46)
    mean, rstd = rand(C, seed=47) * 0.1, rand(C, seed=48).abs() + 0.5
    y = ops().affine_act(z, scale, shift, "tanh", None, 0.3, 11)
    y16 = ops().affine_act(z, scale, shift, "tanh", None, 0.3, 11, out_bf16=True)
    assert torch.equal(y16, y.to(torch.bfloat16))
    gz, dg, db = ops().bn_act_bwd(g, z, scale, shift, mean, rstd, "tanh", True, 0.3, 11)
    gz16, dg2, db2 = ops().bn_act_bwd(g, z, scale, shift, mean, rstd, "tanh", True, 0.3, 11, out_bf16=True)
    assert torch.equal(gz16, gz.to(torch.bfloat16)) and torch.equal(dg, dg2) and torch.equal(db, db2)
    assert rel(ops().colsum(gz16), ops().colsum(gz16.float())) <= 1e-6


def test_postnet_bf16_node_follows_the_per_layer_path():
    """The single-node bf16 PostNet (bf16 activations between the layers, TMA-fed convolutions in all three directions)
    against the per-layer 3xTF32 path: forward, input gradient and every parameter gradient."""
    from fastspeech2_lightning_b200 import functional as Fk, ops as o
    from fastspeech2_lightning_b200.fs2.layers import PostNet

    torch.manual_seed(0)
    pn = PostNet(n_mel_channels=80).to(dev()).train()
    pn.dropout_in_training = False
    x = rand(3, 260, 80, seed=49).requires_grad_(True)
    w = rand(3, 260, 80, seed=50)
    res = {}
    for mode in ("tf32x3", "bf16"):
        o.set_precision(mode)
        try:
            for p in pn.parameters():
                p.grad = None
            x.grad = None
            y = Fk.postnet(x, pn, True)
            (y * w).sum().backward()
            res[mode] = (y.detach().clone(), x.grad.clone(), {n: p.grad.clone() for n, p in pn.named_parameters()})
        finally:
            o.set_precision("tf32x3")
    y0, dx0, g0 = res["tf32x3"]
    y1, dx1, g1 = res["bf16"]
    assert rel(y1, y0) <= 3e-2, rel(y1, y0)
    cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm()))
    assert cos(dx1, dx0) >= 0.995, cos(dx1, dx0)
    for n in g0:
        if float(g0[n].norm()) < 1e-4:
            continue  # conv biases in front of a batch-statistic BatchNorm: analytically zero
        assert cos(g1[n], g0[n]) >= 0.99, (n, cos(g1[n], g0[n]))


def test_gemm_bf16_dact_fuses_the_activation_derivative_and_the_dropout_mask():
    M, K, N, p, seed = 700, 256, 1024, 0.2, 5
    g, w = rand(M, K, seed=51), rand(K, N, seed=52, scale=K ** -0.5)
    pre = rand(M, N, seed=53)
    w16, _ = ops().cast_bf16(w)
    pre16, _ = ops().cast_bf16(pre)
    c, c16 = ops().gemm_bf16_dact(g, w16, pre16, "silu", w_mn=True, dropout_p=p, seed=seed, want_c=True)
    dx, _, _ = ops().gemm_bf16(g, w16.reshape(1, K, N), None, w_mn=True)
    u = pre16.float().double()
    sg = torch.sigmoid(u)
    want = dx.double() * (sg * (1 + u * (1 - sg)))
    mask = ops().act_bwd(torch.ones(M, N, device=dev()), None, None, 1.0, None, p, seed)   # keep/(1-p) or 0
    want = (want * mask.double()).float()
    assert rel(c, want) <= 1e-5, rel(c, want)
    assert torch.equal(c16, c.to(torch.bfloat16))


def test_ffn_half_bf16_node_follows_the_per_op_path():
    from fastspeech2_lightning_b200 import functional as Fk, ops as o
    from fastspeech2_lightning_b200.fs2.conformer import _FeedForwardModule

    torch.manual_seed(1)
    ffn = _FeedForwardModule(256, 1024, dropout=0.0).to(dev())
    x = rand(3, 170, 256, seed=54).requires_grad_(True)
    w = rand(3, 170, 256, seed=55)
    res = {}
    for mode in ("tf32x3", "bf16"):
        o.set_precision(mode)
        try:
            for p in ffn.parameters():
                p.grad = None
            x.grad = None
            y = Fk._ffn_half(x, ffn, True, 0.0)
            (y * w).sum().backward()
            res[mode] = (y.detach().clone(), x.grad.clone(), {n: p.grad.clone() for n, p in ffn.named_parameters()})
        finally:
            o.set_precision("tf32x3")
    (y0, dx0, g0), (y1, dx1, g1) = res["tf32x3"], res["bf16"]
    assert rel(y1, y0) <= 1e-2, rel(y1, y0)
    assert rel(dx1, dx0) <= 2e-2, rel(dx1, dx0)
    for n in g0:
        assert rel(g1[n], g0[n]) <= 3e-2, (n, rel(g1[n], g0[n]))
    # with dropout: forward and backward masks agree (the gradient of a zeroed hidden unit is zero): finite-difference-free check
    o.set_precision("bf16")
    try:
        torch.manual_seed(3)
        x2 = rand(2, 130, 256, seed=56).requires_grad_(True)
        y = Fk._ffn_half(x2, ffn, True, 0.3)
        y.sum().backward()
        assert torch.isfinite(x2.grad).all() and float(x2.grad.abs().mean()) > 0
    finally:
        o.set_precision("tf32x3")


@pytest.mark.parametrize("mode", ["bf16", "tf32x3"])
def test_wgrad_l2_reduction_equals_the_workspace_reduce(mode):
    """Single-tap weight gradients fold their split-K partial tiles into dW with red.global.add.v4.f32 (default) or through the
    workspace + reduce kernel (fs2k_wgrad_set_atomic(0)): same partial tiles, only the order of the fp32 additions differs."""
    from fastspeech2_lightning_b200._lib import lib

    g, x = rand(32, 500, 256, seed=51), rand(32, 500, 1024, seed=52)
    fn = (lambda **kw: ops().gemm_wgrad_bf16(g, x, 1, 0, conv_layout=False, **kw)) if mode == "bf16" else \
         (lambda **kw: ops().gemm_wgrad(g, x, 1, 0, False, **kw))
    ops().set_precision(mode)
    try:
        got = fn()
        acc = torch.full_like(got, 2.0)
        fn(accumulate_into=acc)
        lib().fs2k_wgrad_set_atomic(0)
        want = fn()
    finally:
        lib().fs2k_wgrad_set_atomic(1)
        ops().set_precision("tf32x3")
    assert rel(got, want) <= 2e-6, rel(got, want)
    assert rel(acc - 2.0, want) <= 1e-5
