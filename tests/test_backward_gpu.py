"""Gradient parity: every backward kernel against torch.autograd of the same fp32 op on the CPU, and the
whole training step (forward + losses + backward) against the gradients of the unmodified reference
recorded in tests/golden/case_train_*.npz (per-parameter L2 norms and the first 16 entries)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import case_batch, load_case
from test_model_gpu import build_model
from test_ops_gpu import close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def fns():
    from fastspeech2_lightning_b200 import autograd_fns

    return autograd_fns


def leaf(t):
    return t.clone().to(DEV).requires_grad_(True)


def check_grads(outs_gpu, outs_ref, leaves_gpu, leaves_ref, tol=TOL, what="", zero_grad_leaves=()):
    """`zero_grad_leaves`: leaves whose gradient is analytically zero (a bias in front of a training-mode
    BatchNorm): both sides hold cancellation noise, so they are checked against the scale of the inputs."""
    g = torch.Generator().manual_seed(0)
    loss_gpu = loss_ref = 0
    for og, orf in zip(outs_gpu, outs_ref):
        w = torch.randn(orf.shape, generator=g)
        loss_gpu = loss_gpu + (og * w.to(DEV)).sum()
        loss_ref = loss_ref + (orf * w).sum()
    loss_gpu.backward()
    loss_ref.backward()
    for i, (a, b) in enumerate(zip(leaves_gpu, leaves_ref)):
        assert a.grad is not None, f"{what}: no grad for leaf {i}"
        if i in zero_grad_leaves:
            assert float(a.grad.abs().max()) < 1e-3 and float(b.grad.abs().max()) < 1e-3, f"{what}: leaf {i}"
            continue
        close(a.grad, b.grad, tol, f"{what}: grad of leaf {i}")


def test_layernorm_backward():
    g = torch.Generator().manual_seed(1)
    x, w, b = torch.randn(300, 256, generator=g), torch.randn(256, generator=g), torch.randn(256, generator=g)
    r = [t.clone().requires_grad_(True) for t in (x, w, b)]
    l = [leaf(t) for t in (x, w, b)]
    check_grads([fns().layernorm(l[0], l[1], l[2], 1e-5)], [F.layer_norm(r[0], (256,), r[1], r[2], 1e-5)], l, r, what="layernorm")


def test_layernorm_fork_adds_the_skip_gradient_in_the_same_launch():
    """Pre-norm residual block y = x + Linear(LayerNorm(x)): with `layernorm_fork` the skip connection is the node's second
    output, so dL/dx = LN-backward(g_branch) + g_skip comes out of ONE launch (fs2k_layernorm_bwd_add) — same numbers as
    the separate LayerNorm node + autograd's accumulation add, one launch fewer."""
    g = torch.Generator().manual_seed(5)
    x, w, b = torch.randn(3, 70, 256, generator=g), torch.randn(256, generator=g), torch.randn(256, generator=g)
    lw, lb = torch.randn(256, 256, generator=g) / 16, torch.randn(256, generator=g)
    r = [t.clone().requires_grad_(True) for t in (x, w, b, lw, lb)]
    l = [leaf(t) for t in (x, w, b, lw, lb)]
    ref = r[0] + F.linear(F.layer_norm(r[0], (256,), r[1], r[2], 1e-5), r[3], r[4])
    ln, skip = fns().layernorm_fork(l[0], l[1], l[2], 1e-5)
    assert skip.data_ptr() == l[0].data_ptr()          # no copy of the residual stream
    out = fns().linear(ln, l[3], l[4], None, 1.0, skip)
    check_grads([out], [ref], l, r, tol=2e-4, what="layernorm fork")
    # the skip output alone, and the norm output alone, still back-propagate
    l2 = leaf(x)
    _, s2 = fns().layernorm_fork(l2, l[1].detach().requires_grad_(True), l[2].detach().requires_grad_(True), 1e-5)
    (s2 * 2.0).sum().backward()
    assert torch.equal(l2.grad, torch.full_like(l2, 2.0))


@pytest.mark.parametrize("B,L,K,N,taps,act,res", [(2, 70, 256, 1024, 1, "silu", False), (3, 50, 1024, 256, 1, None, True),
                                                  (2, 90, 80, 512, 5, None, False), (2, 40, 256, 512, 3, "relu", False),
                                                  (2, 33, 512, 80, 5, None, False), (1, 200, 256, 768, 1, None, False)])
def test_gemm_backward(B, L, K, N, taps, act, res):
    g = torch.Generator().manual_seed(L)
    x = torch.randn(B, L, K, generator=g)
    w = torch.randn(N, K, taps, generator=g) / (K * taps) ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    rs = torch.randn(B, L, N, generator=g)
    r = [t.clone().requires_grad_(True) for t in (x, w, b, rs)]
    l = [leaf(t) for t in (x, w, b, rs)]
    ref = F.conv1d(r[0].transpose(1, 2), r[1], r[2], padding=(taps - 1) // 2).transpose(1, 2)
    if act:
        ref = {"silu": F.silu, "relu": F.relu}[act](ref)
    if res:
        ref = ref * 0.5 + r[3]
        got = fns().linear(l[0], l[1].squeeze(-1), l[2], act, 0.5, l[3])
    elif taps == 1:
        got = fns().linear(l[0], l[1].squeeze(-1), l[2], act, 1.0, None)
    else:
        got = fns().conv1d(l[0], l[1], l[2], act)
    n = 4 if res else 3
    check_grads([got], [ref], l[:n], r[:n], what=f"gemm {K}->{N} taps{taps} {act}")


@pytest.mark.parametrize("B,L,lens", [(2, 100, [100, 37]), (3, 130, [130, 64, 1])])
def test_attention_backward(B, L, lens):
    H, hd = 2, 128
    D = H * hd
    g = torch.Generator().manual_seed(L)
    qkv = torch.randn(B, L, 3 * D, generator=g)
    lens_t = torch.tensor(lens, dtype=torch.int32)
    r, l = qkv.clone().requires_grad_(True), leaf(qkv)
    q, k, v = [t.view(B, L, H, hd).transpose(1, 2) for t in r.split(D, -1)]
    s = q @ k.transpose(-1, -2) / hd ** 0.5
    s = s.masked_fill((torch.arange(L)[None, :] >= lens_t[:, None])[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, L, D)
    check_grads([fns().attention(l, lens_t.to(DEV), H)], [ref], [l], [r], what="attention")


@pytest.mark.parametrize("training", [True, False])
def test_conv_module_middle_backward(training):
    g = torch.Generator().manual_seed(3)
    B, L, C, K = 3, 60, 256, 9
    h = torch.randn(B, L, 2 * C, generator=g)
    w, b = torch.randn(C, 1, K, generator=g) / 3, torch.randn(C, generator=g) * 0.1
    bn_ref = torch.nn.BatchNorm1d(C)
    with torch.no_grad():
        bn_ref.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn_ref.bias.copy_(torch.randn(C, generator=g) * 0.1)
        bn_ref.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    bn_gpu = torch.nn.BatchNorm1d(C)
    bn_gpu.load_state_dict(bn_ref.state_dict())
    bn_gpu = bn_gpu.to(DEV)
    bn_ref.train(training)
    r = [t.clone().requires_grad_(True) for t in (h, w, b)]
    l = [leaf(t) for t in (h, w, b)]
    z = F.conv1d(F.glu(r[0], -1).transpose(1, 2), r[1], r[2], padding=4, groups=C)
    ref = F.silu(bn_ref(z)).transpose(1, 2)
    got = fns().glu_dwconv_bn_silu(l[0], l[1], l[2], bn_gpu, training)
    check_grads([got], [ref], l + [bn_gpu.weight, bn_gpu.bias], r + [bn_ref.weight, bn_ref.bias], what="conv module",
                zero_grad_leaves=(2,) if training else ())


def test_postnet_block_backward():
    g = torch.Generator().manual_seed(4)
    B, L, Ci, Co = 2, 50, 80, 512
    x = torch.randn(B, L, Ci, generator=g)
    w, b = torch.randn(Co, Ci, 5, generator=g) / 20, torch.randn(Co, generator=g) * 0.1
    bn_ref = torch.nn.BatchNorm1d(Co)
    bn_gpu = torch.nn.BatchNorm1d(Co).to(DEV)
    r = [t.clone().requires_grad_(True) for t in (x, w, b)]
    l = [leaf(t) for t in (x, w, b)]
    ref = torch.tanh(bn_ref(F.conv1d(r[0].transpose(1, 2), r[1], r[2], padding=2))).transpose(1, 2)
    got = fns().conv1d_bn_act(l[0], l[1], l[2], bn_gpu, "tanh", True)
    # (weight gradients are accumulated with fp32 atomics: run-to-run order noise of a few 1e-5 relative)
    check_grads([got], [ref], l + [bn_gpu.weight, bn_gpu.bias], r + [bn_ref.weight, bn_ref.bias], tol=3e-4,
                what="postnet block", zero_grad_leaves=(2,))


def test_small_ops_backward():
    f = fns()
    g = torch.Generator().manual_seed(5)
    # depthwise k=3 (+bias)
    x, w, b = torch.randn(2, 40, 256, generator=g), torch.randn(256, 1, 3, generator=g), torch.randn(256, generator=g)
    r = [t.clone().requires_grad_(True) for t in (x, w, b)]
    l = [leaf(t) for t in (x, w, b)]
    check_grads([f.dwconv(l[0], l[1], l[2])], [F.conv1d(r[0].transpose(1, 2), r[1], r[2], padding=1, groups=256).transpose(1, 2)], l, r, what="dwconv3")
    # Linear(D→1)·mask
    x, w, b = torch.randn(3, 30, 256, generator=g), torch.randn(1, 256, generator=g) / 16, torch.randn(1, generator=g)
    mask = torch.rand(3, 30, generator=g) > 0.3
    r = [t.clone().requires_grad_(True) for t in (x, w, b)]
    l = [leaf(t) for t in (x, w, b)]
    check_grads([f.rowdot(l[0], l[1], l[2], mask.to(DEV))], [F.linear(r[0], r[1], r[2]).squeeze(-1) * mask], l, r, what="rowdot")
    # masked losses
    p, t = torch.randn(3, 30, 80, generator=g), torch.randn(3, 30, 80, generator=g)
    for kind, fn in (("mse", F.mse_loss), ("mae", F.l1_loss)):
        r, l = p.clone().requires_grad_(True), leaf(p)
        check_grads([f.masked_loss(l, t.to(DEV), mask.to(DEV), kind, 0.7)], [fn(r * mask[..., None], t * mask[..., None]) * 0.7], [l], [r], what=kind)
    # binarisation loss
    soft = torch.softmax(torch.randn(2, 1, 40, 12, generator=g), -1)
    hard = F.one_hot(torch.randint(0, 12, (2, 1, 40), generator=g), 12).float()
    r, l = soft.clone().requires_grad_(True), leaf(soft)
    ref = -torch.log(torch.clamp(r[hard == 1], min=1e-12)).sum() / hard.sum()
    check_grads([f.attn_bin_loss(hard.to(DEV), l)], [ref], [l], [r], what="bin loss")


def test_length_regulator_and_embedding_backward():
    from fastspeech2_lightning_b200 import ops
    from oracle import fs2_oracle

    f = fns()
    g = torch.Generator().manual_seed(6)
    B, T, D = 3, 17, 256
    x = torch.randn(B, T, D, generator=g)
    dur = torch.randint(0, 6, (B, T), generator=g, dtype=torch.int32)
    inv_freq = 1 / (10000 ** (torch.arange(0.0, D, 2.0) / D))
    r, l = x.clone().requires_grad_(True), leaf(x)
    cum, total = ops.lr_scan(dur.to(DEV))
    width = int(total.max())
    out, out_pos, mask = f.length_regulate(l, cum, total, width, inv_freq.to(DEV))
    ref_out, _ = fs2_oracle.length_regulator(r, dur, width)
    check_grads([out, out_pos], [ref_out, ref_out * 1.0], [l], [r], what="length regulator")
    # token embedding (+ posenc): pad row gets no gradient
    table = torch.randn(30, D, generator=g)
    text = torch.randint(0, 30, (B, T), generator=g, dtype=torch.int32)
    lens = torch.tensor([17, 9, 13], dtype=torch.int32)
    r, l = table.clone().requires_grad_(True), leaf(table)
    emb, xx = f.embed_posenc(text.to(DEV), l, inv_freq.to(DEV), lens.to(DEV), 0)
    e_ref = F.embedding(text.long(), r, padding_idx=0)
    check_grads([emb, xx], [e_ref, e_ref * 1.0], [l], [r], what="embed")
    # bucketize + embedding add
    tab, xin = torch.randn(256, D, generator=g), torch.randn(B, T, D, generator=g)
    v = torch.randn(B, T, generator=g) * 2
    bins = torch.linspace(-3, 3, 255)
    r = [t.clone().requires_grad_(True) for t in (xin, tab)]
    l = [leaf(t) for t in (xin, tab)]
    y, ids = f.bucketize_embed_add(v.to(DEV), 1.0, bins.to(DEV), l[1], l[0])
    check_grads([y], [r[0] + F.embedding(torch.bucketize(v, bins), r[1])], l, r, what="bucketize embed add")
    # broadcast rows
    rows, ids = torch.randn(5, D, generator=g), torch.tensor([4, 0, 4], dtype=torch.int32)
    r = [t.clone().requires_grad_(True) for t in (xin, rows)]
    l = [leaf(t) for t in (xin, rows)]
    check_grads([f.add_rows(l[0], [(l[1], ids.to(DEV))])], [r[0] + r[1][ids.long()][:, None]], l, r, what="add_rows")


def test_aligner_backward():
    g = torch.Generator().manual_seed(7)
    B, F_, T = 2, 90, 21
    q, k = torch.randn(B, F_, 80, generator=g) * 3, torch.randn(B, T, 80, generator=g) * 3
    prior = torch.rand(B, F_, T, generator=g)
    lens = torch.tensor([21, 13], dtype=torch.int32)
    r = [t.clone().requires_grad_(True) for t in (q, k)]
    l = [leaf(t) for t in (q, k)]
    d = -0.0005 * ((r[0][:, :, None] - r[1][:, None]) ** 2).sum(-1)
    lp = torch.log_softmax(d, -1) + torch.log(prior + 1e-8)
    pad = torch.arange(T)[None, :] >= lens[:, None]
    masked = lp.clone()
    masked.data.masked_fill_(pad[:, None, :], float("-inf"))
    soft_ref = torch.softmax(masked, -1)
    soft, logprob = fns().aligner_scores(l[0], l[1], prior.to(DEV), lens.to(DEV))
    check_grads([soft[:, 0], logprob[:, 0]], [soft_ref, lp], l, r, what="aligner")


@pytest.mark.parametrize("bwd", [None])  # single-pass TF32 gradients drift to ~1e-2 in the first layers: not used
@pytest.mark.parametrize("name", ["train_bn", "train_frame_level"])
def test_training_step_gradients_match_reference(name, bwd, request):
    from fastspeech2_lightning_b200 import ops

    ops.set_precision("tf32x3", bwd)
    request.addfinalizer(lambda: ops.set_precision("tf32x3", None))
    meta, gold = load_case(name)
    model = build_model(meta)
    batch = case_batch(meta, DEV)
    out = model(batch)
    losses = model.loss(out, batch, model.current_epoch)
    for k, want in gold.items():
        if k.startswith("loss."):
            close(losses[k[5:]].detach().cpu().reshape(()), np.asarray(want, dtype=np.float64), TOL, f"{name}:{k}")
    losses["total"].backward()
    params = dict(model.named_parameters())
    worst = 0.0
    for n, norm, head in zip([str(x) for x in gold["grad.names"]], gold["grad.norms"], gold["grad.heads"]):
        gr = params[n].grad
        assert gr is not None, f"no gradient for {n}"
        gr = gr.detach().double().flatten().cpu()
        if norm < 1e-6:
            # analytically zero (a bias in front of a training-mode BatchNorm): both sides are rounding noise
            assert float(gr.norm()) < 1e-5, n
            continue
        rel = abs(float(gr.norm()) - norm) / max(norm, 1e-6)
        worst = max(worst, rel)
        print(f"  {rel:.2e} {n}") if rel > 1e-3 else None
        # 3×TF32 forward and backward: measured worst case 2.7e-5; 2e-4 leaves headroom for a different
        # summation order, not for a wrong term.  Single-pass TF32 gradients (the `bwd` variants) carry ≈ 1e-3 rounding.
        assert rel <= (2e-4 if bwd is None else 1e-2), f"{n}: |grad| {float(gr.norm()):.6e} vs reference {norm:.6e}"
        h = torch.zeros(16, dtype=torch.float64)
        h[: min(16, gr.numel())] = gr[:16]
        scale = max(float(np.abs(head).max()), norm / max(gr.numel(), 1) ** 0.5, 1e-9)
        assert float((h - torch.from_numpy(head).double()).abs().max()) <= (5e-4 if bwd is None else 3e-2) * scale + 1e-7, n
    for k, want in gold.items():
        if k.startswith("bn."):
            close(model.state_dict()[k[3:]], want, 1e-4, k)
    print(f"{name}: worst relative gradient-norm error {worst:.2e} over {len(gold['grad.norms'])} parameters")


@pytest.mark.parametrize("B,F,T,case", [(4, 60, 12, "ragged"), (3, 33, 40, "impossible"), (32, 500, 80, "c2"), (2, 700, 300, "long")])
def test_ctc_forward_sum_kernel_matches_torch_ctc(B, F, T, case):
    """csrc/ctc.cu against the reference composition (pad blank, mask keys, log_softmax, nn.CTCLoss) evaluated
    by the oracle on the CPU: loss and gradient w.r.t. attn_logprob.  'impossible': query_len < key_len for one
    utterance → infinite nll → zero_infinity zeroes that utterance's loss and gradient."""
    from fastspeech2_lightning_b200 import autograd_fns as fns
    from oracle.fs2_oracle import attention_ctc_loss

    g = torch.Generator().manual_seed(B * 1000 + F)
    # the oracle composition evaluated in fp64 is the truth; torch's own fp32 recursion carries ≈1e-3 of rounding
    # noise on the posteriors at F = 500 (|α| ≈ 1500, fp32 spacing 1e-4), the kernel recurses in fp64
    x = (torch.randn(B, 1, F, T, generator=g) * 2.0).double().requires_grad_(True)
    key_lens = torch.randint(max(T // 2, 1), T + 1, (B,), generator=g)
    query_lens = torch.randint(max(F // 2, T), F + 1, (B,), generator=g) if case != "impossible" else torch.tensor([F, 10, F])
    key_lens[0], query_lens[0] = T, F
    if case == "impossible":
        key_lens[1] = 30  # 10 frames cannot emit 30 tokens
    ref = attention_ctc_loss(x, key_lens, query_lens)
    (ref * 1.7).backward()
    xd = x.detach().float().to(DEV).requires_grad_(True)
    out = fns.ctc_forward_sum(xd, key_lens.to(DEV), query_lens.to(DEV), -1.0)
    (out * 1.7).backward()
    assert abs(float(out.detach()) - float(ref.detach())) <= 2e-6 * max(1.0, abs(float(ref.detach()))), (float(out.detach()), float(ref.detach()))
    gref, ggpu = x.grad, xd.grad.cpu().double()
    err = float((gref - ggpu).abs().max()) / max(float(gref.abs().max()), 1e-30)
    assert err <= 2e-6, err
    if case == "c2":  # and the reference's own fp32 arithmetic agrees within its noise
        x32 = x.detach().float().requires_grad_(True)
        r32 = attention_ctc_loss(x32, key_lens, query_lens)
        (r32 * 1.7).backward()
        assert abs(float(r32.detach()) - float(out.detach())) <= 1e-4 * abs(float(r32.detach()))
        assert float((x32.grad.double() - ggpu).abs().max()) / float(gref.abs().max()) <= 1e-2
    # gradient past the utterance and for masked keys is exactly zero
    for b in range(B):
        assert float(ggpu[b, 0, int(query_lens[b]):].abs().max() if int(query_lens[b]) < F else 0.0) == 0.0
        assert float(ggpu[b, 0, :, int(key_lens[b]):].abs().max() if int(key_lens[b]) < T else 0.0) == 0.0
    if case == "impossible":
        assert float(ggpu[1].abs().max()) == 0.0


@pytest.mark.parametrize("B,F", [(4, 130), (8, 500)])
def test_gst_style_encoder_training_kernels_match_fp64_autograd(B, F):
    """Training through the GST StyleEncoder on libfs2k kernels (raw conv + BatchNorm batch statistics + ReLU, GRU with
    BPTT, token attention) against torch autograd over the same module in float64 on the CPU (the GPU library path runs
    its convolutions in TF32 by default and is itself 1e-2 away): output, every parameter gradient, BatchNorm running
    statistics."""
    import copy

    from fastspeech2_lightning_b200.fs2.gst import model as gst
    from helpers import gst_reference_forward

    torch.manual_seed(F)
    ours = gst.StyleEncoder(idim=80).train()
    ref = copy.deepcopy(ours).double()
    ours = ours.to(DEV)
    speech = torch.randn(B, F, 80)
    go = torch.randn(B, 256)
    y_ref = gst_reference_forward(ref, speech.double())  # plain torch over the module's own (fp64, CPU) parameters
    y_ref.backward(go.double())
    y = ours(speech.to(DEV))
    y.backward(go.to(DEV))
    close(y.detach(), y_ref.detach(), 1e-4, "style embedding (training mode)")
    for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, n
        scale = float(q.grad.abs().max())
        err = float((p.grad.cpu().double() - q.grad).abs().max())
        assert err <= 1e-3 * max(scale, 1e-9) + 1e-7, (n, err, scale)
    for (n, b1), (_, b2) in zip(ours.named_buffers(), ref.named_buffers()):
        if b1.is_floating_point():
            close(b1, b2, 1e-4, n)
        else:
            assert torch.equal(b1.cpu(), b2), n
