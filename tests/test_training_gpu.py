"""Training-step tail: flat-buffer AdamW (+ fused clipping) against torch.optim.AdamW + clip_grad_norm_,
dropout statistics, and a short optimisation run of the whole model."""
import pytest
import torch

from helpers import case_batch, load_case
from test_model_gpu import build_model
from test_ops_gpu import close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_fused_adamw_matches_torch_adamw_with_clipping():
    from fastspeech2_lightning_b200.optim import FusedAdamW

    g = torch.Generator().manual_seed(0)
    shapes = [(256, 1024), (1024,), (80, 256, 5), (3,), (1, 256)]
    ours = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    kw = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01)
    opt = FusedAdamW(ours, max_grad_norm=1.0, **kw)
    opt_ref = torch.optim.AdamW(ref, **kw)
    for step in range(4):
        grads = [torch.randn(s, generator=g) * (5.0 if step % 2 == 0 else 0.01) for s in shapes]
        opt.zero_grad()
        for p, r, gr in zip(ours, ref, grads):
            p.grad.copy_(gr.to(DEV))
            r.grad = gr.to(DEV).clone()
        total = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        close(opt.grad_norm().reshape(()), total.reshape(()), 1e-5, "grad norm")
        opt.step()
        opt_ref.step()
        for p, r in zip(ours, ref):
            close(p, r, 2e-6, f"param after step {step}")
    assert all(p.grad.data_ptr() >= opt.flat_g.data_ptr() for p in ours)


def test_dropout_statistics_and_backward_mask():
    from fastspeech2_lightning_b200 import autograd_fns as fns

    x = torch.ones(1 << 20, device=DEV, requires_grad=True)
    y = fns.dropout(x, 0.3)
    keep = (y != 0).float().mean().item()
    assert abs(keep - 0.7) < 5e-3
    assert torch.allclose(y[y != 0], torch.full((1,), 1 / 0.7, device=DEV))
    y.sum().backward()
    assert torch.equal(x.grad != 0, y != 0)  # the backward regenerates the same mask


def test_attention_dropout_forward_backward_consistency():
    """With dropout the attention gradient must be the gradient of the same masked computation: check
    against finite differences of the kernel itself (the mask is a pure function of the seed)."""
    from fastspeech2_lightning_b200 import ops

    g = torch.Generator().manual_seed(3)
    B, L, H, hd = 1, 40, 2, 128
    qkv = (torch.randn(B, L, 3 * H * hd, generator=g) * 0.5).to(DEV)
    lens = torch.tensor([33], dtype=torch.int32, device=DEV)
    w = torch.randn(B, L, H * hd, generator=g).to(DEV)
    out, lse = ops.attention(qkv, lens, H, want_lse=True, dropout_p=0.25, seed=77)
    dqkv = ops.attention_bwd(qkv, out, lse, w, lens, H, 0.25, 77)
    out0 = ops.attention(qkv, lens, H, dropout_p=0.0)
    assert (out - out0).abs().max() > 1e-3  # dropout is active
    for idx in [(0, 3, 5), (0, 20, 256 + 17), (0, 10, 512 + 200), (0, 32, 700)]:
        e = torch.zeros_like(qkv)
        e[idx] = 1e-2
        fp = (ops.attention(qkv + e, lens, H, dropout_p=0.25, seed=77) * w).sum()
        fm = (ops.attention(qkv - e, lens, H, dropout_p=0.25, seed=77) * w).sum()
        fd = float(fp - fm) / 2e-2
        assert abs(fd - float(dqkv[idx])) <= 2e-2 * max(1.0, abs(fd)), (idx, fd, float(dqkv[idx]))


def test_training_run_reduces_the_loss():
    meta, _ = load_case("train_bn")
    model = build_model(meta)
    model.fused_grad_clip = 1.0
    model.postnet.dropout_in_training = True
    (opt,), (sched,) = model.configure_optimizers()
    batch = case_batch(meta, DEV)
    first = last = None
    for i in range(12):
        opt.zero_grad()
        loss = model.training_step(batch, i)
        loss.backward()
        opt.step()
        sched["scheduler"].step()
        first = float(loss) if first is None else first
        last = float(loss)
    assert all(torch.isfinite(p).all() for p in model.parameters())
    assert last < first, (first, last)
    assert "training/total_loss" in model.logged
