"""CPU tests of host-side logic that needs no kernel: the opt-in bucketed padding, the two-phase backward the data-parallel
step is built on (autograd's `inputs=` pruning around a detached cut), and the graph-capture back-off rule."""
import types

import torch

from fastspeech2_lightning_b200 import synthetic
from fastspeech2_lightning_b200.fs2.batching import pad_batch_to_multiple
from fastspeech2_lightning_b200.graphs import GraphedTrainStep


def test_pad_batch_to_multiple_rounds_shapes_up_and_keeps_everything_valid():
    b = synthetic.make_batch(4, (60, 80), seed=5, learn_alignment=True)     # T = 80 == n_mels on purpose
    T, F = int(b["max_src_len"]), int(b["max_mel_len"])
    p = pad_batch_to_multiple(b, (32, 32))
    T2, F2 = int(p["max_src_len"]), int(p["max_mel_len"])
    assert (T2, F2) == (96, (F + 31) // 32 * 32)
    assert p["mel"].shape == (4, F2, 80)                                      # the channel dimension is not a length
    assert p["duration"].shape == (4, F2, T2) and p["text"].shape == (4, T2) and p["pitch"].shape == (4, F2)
    assert torch.equal(p["mel"][:, :F], b["mel"]) and float(p["mel"][:, F:].abs().sum()) == 0.0
    assert torch.equal(p["duration"][:, :F, :T], b["duration"]) and float(p["duration"][:, F:].abs().sum()) == 0.0
    assert torch.equal(p["src_lens"], b["src_lens"]) and torch.equal(p["mel_lens"], b["mel_lens"])
    assert pad_batch_to_multiple(p, (32, 32)) is p                            # already on the grid: untouched
    q = pad_batch_to_multiple(synthetic.make_batch(2, (10, 14), seed=1, learn_alignment=False, inference=True), (8, 32))
    assert q["text"].shape[1] == 16                                           # inference batches: only the text side exists


class _Probe(torch.autograd.Function):
    calls = []

    @staticmethod
    def forward(ctx, x, w, name):
        ctx.name = name
        ctx.save_for_backward(x, w)
        return x * w

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        _Probe.calls.append(ctx.name)
        return g * w, g * x, None


def test_two_phase_backward_around_a_cut_runs_every_node_once_and_equals_the_single_pass():
    """graphs.GraphedTrainStep: phase 1 = total.backward(inputs=late-stage parameters + the cut leaf); phase 2 =
    backward([the cut's producer, total], [leaf.grad, None], inputs=the other parameters).  Losses that bypass the cut
    (variance / aligner losses) are back-propagated in phase 2 only, nothing runs twice, the gradients are the single-pass ones."""
    torch.manual_seed(0)
    a, b, c = (torch.randn(5, requires_grad=True) for _ in range(3))

    def graph(cut):
        h = _Probe.apply(a, b, "encoder")
        l_side = (h ** 2).sum()                      # a loss that does not pass through the decoder
        up = h * 2
        leaf = up.detach().requires_grad_(True) if cut else up
        y = _Probe.apply(leaf, c, "decoder")
        return up, leaf, y.sum() + 0.5 * l_side

    _Probe.calls.clear()
    up, leaf, total = graph(cut=True)
    total.backward(inputs=[c, leaf], retain_graph=True)
    assert _Probe.calls == ["decoder"] and a.grad is None and b.grad is None and c.grad is not None
    torch.autograd.backward([up, total], [leaf.grad, None], inputs=[a, b])
    assert _Probe.calls == ["decoder", "encoder"]
    two_phase = [t.grad.clone() for t in (a, b, c)]
    for t in (a, b, c):
        t.grad = None
    graph(cut=False)[2].backward()
    for got, want in zip(two_phase, (a.grad, b.grad, c.grad)):
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)


def test_capture_backs_off_when_captured_shapes_do_not_recur():
    pays = GraphedTrainStep._capture_pays
    r = types.SimpleNamespace(_n_captures=0, _n_replays=0)
    assert pays(r)                                   # the first captures are free
    r._n_captures = 4
    assert not pays(r)                               # four captures and no replay: the stream is shape-diverse, stay eager
    r._n_replays = 15
    assert not pays(r)
    r._n_replays = 16
    assert pays(r)                                   # four replays per capture: capturing pays again
    r._n_captures, r._n_replays = 18, 500
    assert pays(r)
