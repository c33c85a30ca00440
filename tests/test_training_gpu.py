"""Training-step tail: flat-buffer AdamW (+ fused clipping) against torch.optim.AdamW + clip_grad_norm_,
dropout statistics, and a short optimisation run of the whole model."""
import numpy as np
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


def _fresh_model(meta, clip=1.0):
    model = build_model(meta)
    model.fused_grad_clip = clip
    model.configure_optimizers()
    return model


def test_graph_replayed_training_step_matches_eager_steps():
    """graphs.GraphedTrainStep: first sight eager, then capture + replay.  Same batches, dropout off → the loss
    trajectory and the parameters follow the eager run (differences only from the fp64 atomics' summation order)."""
    meta, _ = load_case("train_bn")
    batch = case_batch(meta, DEV)
    runs = {}
    for mode in ("eager", "graph"):
        torch.manual_seed(0)
        model = _fresh_model(meta)
        traj = []
        for i in range(5):
            losses = model.optimization_step(batch, use_cuda_graph=(mode == "graph"))
            traj.append({k: float(v) for k, v in losses.items()})
        runs[mode] = (traj, [p.detach().clone() for p in model.parameters()], model)
    assert len(runs["graph"][2]._train_runner._cache) == 1  # two eager sights, captured once, replayed three times
    for a, b in zip(runs["eager"][0], runs["graph"][0]):
        for k in a:
            assert abs(a[k] - b[k]) <= 2e-4 * max(1.0, abs(a[k])), (k, a[k], b[k])
    assert runs["eager"][0][-1]["total"] < runs["eager"][0][0]["total"]
    worst = max(float((p - q).abs().max()) for p, q in zip(runs["eager"][1], runs["graph"][1]))
    assert worst < 5e-4, worst
    assert runs["graph"][2].optimizer._step == 5 and runs["graph"][2].scheduler.last_epoch == 5


def test_graph_replay_draws_fresh_dropout_masks_and_learning_rates():
    """By-value seeds are frozen inside a captured launch; the per-step seed base in device memory must still give
    every replay its own masks, and the learning rate must come from device memory (Noam moves it every step)."""
    meta, _ = load_case("train_bn")
    batch = case_batch(meta, DEV)
    model = _fresh_model(meta)
    model.postnet.dropout_in_training = True
    for g in model.optimizer.param_groups:
        g["weight_decay"] = 0.0
    model.scheduler.base_lrs = [0.0 for _ in model.scheduler.base_lrs]  # frozen parameters: only dropout varies
    seen = []
    for i in range(6):  # two eager sights, capture + replay, three more replays
        losses = model.optimization_step(batch)
        seen.append({k: float(v) for k, v in losses.items()})
    assert len(model._train_runner._cache) == 1
    replays = seen[2:]
    assert len({round(r["postnet"], 9) for r in replays}) == len(replays), replays  # fresh PostNet dropout masks
    assert all(abs(r["spec"] - replays[0]["spec"]) < 1e-6 for r in replays)          # no dropout before the PostNet
    # now let the schedule drive the captured optimizer: parameters must move by the current lr, not the captured one
    model.scheduler.base_lrs = [1e-3 for _ in model.scheduler.base_lrs]
    before = model.optimizer.flat_p.clone()
    model.scheduler.step()
    lr_now = model.optimizer.param_groups[0]["lr"]
    assert lr_now > 0
    model.optimization_step(batch)
    moved = float((model.optimizer.flat_p - before).abs().max())
    assert 0 < moved <= 1.01 * lr_now * 1.5, (moved, lr_now)  # |Δp| ≤ lr·|m̂/√v̂| — bounded by ≈ lr early in training


def test_side_stream_weight_gradients_equal_autograd_accumulated_ones():
    """WgradSink: wgrad + bias column sums on a second stream, accumulated straight into the flat gradient buffer,
    must leave exactly the gradient autograd's AccumulateGrad path leaves."""
    from fastspeech2_lightning_b200.graphs import GraphedTrainStep

    meta, _ = load_case("train_bn")
    batch = case_batch(meta, DEV)
    flat = {}
    for overlap in (False, True):
        model = _fresh_model(meta, clip=None)
        for g in model.optimizer.param_groups:
            g["lr"] = 0.0
        model._train_runner = GraphedTrainStep(model, model.optimizer, None, overlap_wgrad=overlap)
        model._train_runner._step_body(batch)
        torch.cuda.synchronize()
        flat[overlap] = model.optimizer.flat_g.clone()
    assert float(flat[False].abs().max()) > 0
    diff = (flat[False] - flat[True]).abs()
    scale = float(flat[False].abs().max())
    assert float(diff.max()) <= 1e-6 * scale, (float(diff.max()), scale, int((diff > 0).sum()))


def test_gemm_epilogue_dropout_equals_separate_dropout_kernel():
    """residual + Dropout(x·Wᵀ + b)·alpha fused in the tensor-core epilogue == GEMM then fs2k_dropout (same seed →
    same counter-hash mask), and the autograd Function's gradients match the unfused composition."""
    from fastspeech2_lightning_b200 import autograd_fns as fns
    from fastspeech2_lightning_b200 import ops

    g = torch.Generator().manual_seed(5)
    B, L, K, N = 3, 77, 256, 256
    x = torch.randn(B, L, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / 16).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(B, L, N, generator=g).to(DEV)
    fused = ops.gemm(x, w, b, alpha=0.5, residual=res, dropout_p=0.2, seed=1234)
    plain = ops.gemm(x, w, b, alpha=0.5)
    ref = ops.dropout(plain, 0.2, 1234, res)
    assert torch.equal(fused, ref)
    assert abs(float((fused == res).float().mean()) - 0.2) < 0.02  # ≈ 20 % of the values were dropped

    # gradients through the Function (seed drawn from torch's CPU generator: reseed to replay it)
    xs = [x.clone().requires_grad_(True) for _ in range(2)]
    ws = [w.clone().requires_grad_(True) for _ in range(2)]
    rs = [res.clone().requires_grad_(True) for _ in range(2)]
    go = torch.randn(B, L, N, generator=g).to(DEV)
    torch.manual_seed(11)
    y0 = fns.linear(xs[0], ws[0], b, None, 0.5, rs[0], dropout_p=0.2)
    torch.manual_seed(11)
    y1 = fns.dropout(fns.linear(xs[1], ws[1], b, None, 0.5, None), 0.2, rs[1])
    assert torch.equal(y0, y1)
    y0.backward(go)
    y1.backward(go)
    close(xs[0].grad, xs[1].grad, 1e-5, "dx")
    close(ws[0].grad, ws[1].grad, 1e-5, "dw")
    assert torch.equal(rs[0].grad, rs[1].grad)


def test_optimization_step_with_gst_reference_encoder_trains():
    """Training through the GST reference encoder uses torch autograd over library kernels (conv2d / GRU): the step must
    still run through `optimization_step` — captured if the library ops allow it, with eager launches otherwise."""
    import warnings

    meta, _ = load_case("train_gst")
    batch = case_batch(meta, DEV)
    model = _fresh_model(meta)
    model.train()  # cuDNN's RNN backward exists in training mode only
    before = model.optimizer.flat_p.clone()
    losses = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(4):
            losses.append(float(model.optimization_step(batch)["total"]))
    assert all(np.isfinite(v) for v in losses), losses  # (dropout is on: four steps are too noisy to demand a decrease)
    assert model.optimizer._step == 4 and bool(torch.isfinite(model.optimizer.flat_p).all())
    gst = [p for n, p in model.named_parameters() if n.startswith("gst.ref_enc")]
    assert float((model.optimizer.flat_p - before).abs().max()) > 0 and all(float(p.grad.abs().max()) > 0 for p in gst)


@pytest.mark.parametrize("graph", [False, True])
def test_repacked_conv_weights_follow_the_optimizer(graph):
    """The update kernel writes parameters through raw pointers; everything cached per parameter version (the
    [N,K,taps] → [taps,N,K] re-pack of Conv1d weights) must still see the new values — eagerly and inside replays."""
    from fastspeech2_lightning_b200 import ops

    meta, _ = load_case("train_bn")
    batch = case_batch(meta, DEV)
    model = _fresh_model(meta)
    model.scheduler.base_lrs = [1.0 for _ in model.scheduler.base_lrs]  # Noam: lr = 2.5e-4·step — visible updates
    w = next(p for n, p in model.named_parameters() if n.startswith("postnet") and p.dim() == 3 and p.shape[-1] > 1)
    before = w.detach().clone()
    first = ops.conv_weight_taps(w).clone()
    steps = 5 if graph else 2
    for _ in range(steps):
        model.optimization_step(batch, use_cuda_graph=graph)
    assert float((w.detach() - before).abs().max()) > 1e-5  # the optimizer moved it
    now = ops.conv_weight_taps(w)
    assert torch.equal(now, w.detach().permute(2, 0, 1).contiguous()), "stale re-packed conv weight"
    assert not torch.equal(now, first)


def test_three_optimisation_steps_follow_the_oracle_with_torch_adamw():
    """End-to-end training parity over several steps: kernels + FusedAdamW + NoamLR on the GPU against the oracle's
    forward/loss with torch autograd, clip_grad_norm_(1.0), torch.optim.AdamW and the same schedule on the CPU.
    Dropout is off (golden `train_bn` configuration).  A forward that kept using old weights would drift by percents."""
    from fastspeech2_lightning_b200.fs2.noam import NoamLR
    from helpers import case_config, case_state_dict
    from oracle import fs2_oracle

    meta, _ = load_case("train_bn")
    cfg = case_config(meta)
    model = _fresh_model(meta)
    o = cfg.training.optimizer
    sd = {k: v.clone() for k, v in case_state_dict(meta).items()}
    params = []
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var", "_bins", "inv_freq")):
            v.requires_grad_(True)
            params.append(v)
    ref_opt = torch.optim.AdamW(params, lr=o.learning_rate, betas=tuple(o.betas), eps=o.eps, weight_decay=o.weight_decay)
    ref_sched = NoamLR(ref_opt, o.warmup_steps)
    for s in (model.scheduler, ref_sched):
        s.base_lrs = [1.0 for _ in s.base_lrs]  # lr = 2.5e-4·step: three visible updates
    for g, gr in zip(model.optimizer.param_groups, ref_opt.param_groups):
        g["lr"] = gr["lr"] = 2.5e-4
    ocfg = fs2_oracle.Cfg(cfg)
    batch_cpu = case_batch(meta, "cpu")
    batch = case_batch(meta, DEV)
    got, want = [], []
    for step in range(3):
        got.append(float(model.optimization_step(batch, use_cuda_graph=False)["total"]))
        ref_opt.zero_grad()
        out = fs2_oracle.forward(sd, ocfg, batch_cpu, training=True, new_stats={})
        loss = fs2_oracle.loss(out, batch_cpu, ocfg, model.current_epoch)["total"]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        ref_opt.step()
        ref_sched.step()
        want.append(float(loss.detach()))
    print("loss trajectory gpu", got, "oracle", want)
    assert abs(want[2] - want[0]) > 1e-3 * abs(want[0]), "the steps did not change the loss: the test would prove nothing"
    for a, b in zip(got, want):
        assert abs(a - b) <= 2e-4 * abs(b), (got, want)


def test_fused_adamw_checkpoint_is_interchangeable_with_torch_adamw():
    """state_dict / load_state_dict in torch.optim.AdamW's layout: save after two steps, resume in a fresh FusedAdamW AND
    in a torch AdamW, take a third step — all three trajectories agree (moments and bias-correction step restored)."""
    from fastspeech2_lightning_b200.optim import FusedAdamW

    g = torch.Generator().manual_seed(1)
    shapes = [(64, 48), (48,), (16, 8, 3)]
    kw = dict(lr=2e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01)
    grads = [[torch.randn(s, generator=g).to(DEV) for s in shapes] for _ in range(3)]

    def run(opt_cls, params, steps, **extra):
        opt = opt_cls(params, **kw, **extra)
        for gr in steps:
            for p, x in zip(params, gr):
                if p.grad is None:
                    p.grad = x.clone()
                else:
                    p.grad.copy_(x)
            opt.step()
        return opt

    init = [torch.randn(s, generator=g).to(DEV) for s in shapes]
    a = [torch.nn.Parameter(t.clone()) for t in init]
    opt_a = run(FusedAdamW, a, grads[:2])
    sd = opt_a.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 2.0
    # resume: ours from ours
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    opt_b = FusedAdamW(b, **kw)
    opt_b.load_state_dict(sd)
    # resume: torch AdamW from ours
    c = [torch.nn.Parameter(p.detach().clone()) for p in a]
    opt_c = torch.optim.AdamW(c, **kw)
    opt_c.load_state_dict(sd)
    # and ours from torch's own state dict
    d0 = [torch.nn.Parameter(t.clone()) for t in init]
    opt_t = run(torch.optim.AdamW, d0, grads[:2])
    d = [torch.nn.Parameter(p.detach().clone()) for p in d0]
    opt_d = FusedAdamW(d, **kw)
    opt_d.load_state_dict(opt_t.state_dict())
    for params, opt in ((a, opt_a), (b, opt_b), (c, opt_c), (d, opt_d)):
        for p, x in zip(params, grads[2]):
            if p.grad is None:
                p.grad = x.clone()
            else:
                p.grad.copy_(x)
        opt.step()
    for pa, pb, pc, pd in zip(a, b, c, d):
        close(pb, pa, 1e-6, "FusedAdamW resumed from its own checkpoint")
        close(pc, pa, 2e-6, "torch AdamW resumed from a FusedAdamW checkpoint")
        close(pd, pa, 2e-6, "FusedAdamW resumed from a torch AdamW checkpoint")


def test_fused_adamw_adopts_gradients_that_left_the_flat_buffer():
    from fastspeech2_lightning_b200.optim import FusedAdamW

    p = torch.nn.Parameter(torch.ones(40, 8, device=DEV))
    r = torch.nn.Parameter(torch.ones(40, 8, device=DEV))
    opt, ref = FusedAdamW([p], lr=1e-2), torch.optim.AdamW([r], lr=1e-2)
    p.grad = None                      # what nn.Module.zero_grad(set_to_none=True) leaves behind
    (p * 3.0).sum().backward()         # autograd allocates a fresh .grad outside the flat buffer
    (r * 3.0).sum().backward()
    assert p.grad.data_ptr() != opt.flat_g.data_ptr()
    opt.step()
    ref.step()
    close(p, r, 1e-6, "step with a foreign .grad")
    assert p.grad.data_ptr() == opt.flat_g.data_ptr()


@pytest.mark.parametrize("precision", ["tf32x3", "bf16"])
def test_cut_backward_with_overlapped_exchange_equals_the_single_graph_step(precision, monkeypatch):
    """Data-parallel step layout on ONE GPU: the backward cut at the decoder's input, the early bucket's all-reduce on the
    communication stream while the second backward graph runs, then the rest and the update graph.  The all-reduce is
    replaced by "two ranks with identical gradients" (×2 on the stream it is called on), so the result must equal the
    plain one-graph step (tests/test_ddp_gpu.py is the same check over real NCCL with two GPUs)."""
    from fastspeech2_lightning_b200 import ops

    meta, _ = load_case("train_bn")
    batch = case_batch(meta, DEV)
    ops.set_precision(precision)
    try:
        runs = {}
        for mode in ("single", "cut"):
            torch.manual_seed(0)
            model = _fresh_model(meta)
            if mode == "cut":
                calls = []

                def fake_all_reduce(t, group=None, **kw):
                    calls.append((t.numel(), torch.cuda.current_stream().cuda_stream))
                    t.mul_(2)

                monkeypatch.setattr(torch.distributed, "all_reduce", fake_all_reduce)
                monkeypatch.setattr(model.optimizer, "world_size", lambda: 2)
            traj = [float(model.optimization_step(batch, use_cuda_graph=True)["total"]) for _ in range(4)]
            runs[mode] = (traj, model.optimizer.flat_p.detach().clone(), model)
        runner = runs["cut"][2]._train_runner
        lo, hi = runs["cut"][2].optimizer.early_bucket
        assert 0 < lo < hi and len(runner._cache) == 1 and all(len(e) == 5 and e[4] is not None for e in runner._cache.values())   # the three-graph layout was used
        main = torch.cuda.current_stream().cuda_stream
        assert any(n == hi - lo and s != main for n, s in calls), calls       # early bucket went out on the communication stream
        for a, b in zip(runs["single"][0], runs["cut"][0]):
            assert abs(a - b) <= (2e-4 if precision == "tf32x3" else 5e-3) * max(1.0, abs(a)), (runs["single"][0], runs["cut"][0])
        err = float((runs["single"][1] - runs["cut"][1]).abs().max())
        assert err < (5e-4 if precision == "tf32x3" else 2e-3), err
    finally:
        ops.set_precision("tf32x3")


def test_batches_from_the_device_collate_replay_the_captured_step():
    """`fs2.batching.collate_to_device` output fed to `optimization_step`: the private `_staging` blob (whose size depends on
    the valid lengths) is not part of the graph key, so batches with the same padded shape but different lengths hit the same
    captured graph (two eager sights, one capture, then replays), and `_seen` does not grow."""
    from fastspeech2_lightning_b200 import synthetic
    from fastspeech2_lightning_b200.fs2.batching import collate_to_device
    from fastspeech2_lightning_b200.fs2.config import FastSpeech2Config
    from fastspeech2_lightning_b200.fs2.model import FastSpeech2

    def items(seed):
        g = np.random.default_rng(seed)
        out = []
        for b in range(4):
            T = 20 if b == 0 else int(g.integers(8, 20))          # item 0 fixes the padded shape (T = 20, F = 60)
            dur = np.full(T, 3) if b == 0 else g.integers(1, 4, size=T)
            F = int(dur.sum())
            out.append({"text": torch.from_numpy(g.integers(1, 27, size=T).astype(np.int64)),
                        "mel": torch.from_numpy(g.standard_normal((F, 80)).astype(np.float32)),
                        "duration": torch.from_numpy(g.random((F, T)).astype(np.float32)),
                        "pitch": torch.from_numpy(g.standard_normal(F).astype(np.float32)),
                        "energy": torch.from_numpy(g.standard_normal(F).astype(np.float32)),
                        "speaker_id": 0, "language_id": 0, "basename": f"u{seed}-{b}", "duration_control": 1.0,
                        "mel_style_reference": None})
        return out

    model = FastSpeech2(FastSpeech2Config(), stats=synthetic.DEFAULT_STATS)
    synthetic.fill_weights_(model, seed=21)
    model.train()
    model = model.to(DEV)
    model.fused_grad_clip = 1.0
    model.configure_optimizers()
    sizes = set()
    for i in range(5):
        batch = collate_to_device(items(100 + i), DEV, True)
        sizes.add(batch["_staging"].numel())
        loss = float(model.optimization_step(batch)["total"])
        assert loss == loss
    runner = model._train_runner
    assert len(sizes) > 1                                   # the staging blobs did differ …
    assert len(runner._cache) == 1 and len(runner._seen) == 1 and runner._n_replays >= 2   # … and the shape was captured once and replayed
