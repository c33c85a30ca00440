"""CUDA-graph replay of the synthesis forward and of the whole training step.

The model is small (≈200 kernel launches, ≈0.2 TFLOP per batch), so an eager step is bound by host
launch latency, not by the GPU.  `GraphedSynthesis` captures `model(batch, inference=True)` once per
static shape signature (B, T, F) on a side stream and afterwards only copies the new inputs into the
captured input buffers and replays the graph: one launch per step.  Capture needs a sync-free forward,
i.e. given durations (teacher-forced synthesis) with `validate_durations = False`.
"""
from __future__ import annotations

import gc
import itertools
import weakref

import torch

from ._lib import check, lib

_ARMED_SEED_BASE = 0   # device address the library's dropout seed-base pointer currently holds (0 = none)
_ARMED_OWNER = 0       # token of the runner that armed it (the allocator may hand a dead counter's address to a new optimizer)
_owner_tokens = itertools.count(1)


def _clear_seed_base(ptr: int, owner: int) -> None:
    """The counter's owner is gone: a dangling pointer would make every later dropout kernel read freed memory.  Only the
    runner that armed the pointer clears it, and never while a capture is in progress (the reset is a synchronous copy to a
    device symbol — it would invalidate the capture; the next step re-arms anyway)."""
    global _ARMED_SEED_BASE, _ARMED_OWNER
    if _ARMED_SEED_BASE == ptr and _ARMED_OWNER == owner:
        try:
            if not torch.cuda.is_current_stream_capturing():
                lib().fs2k_set_dropout_seed_base(None)
        except Exception:
            pass
        _ARMED_SEED_BASE = _ARMED_OWNER = 0


def _signature(batch) -> tuple:
    """Shape signature of the model's input keys.  Private keys (`_staging`, the packed host blob collate_to_device adds,
    whose size depends on the valid lengths) are not inputs of the model and stay out of the key — and out of the
    captured input buffers."""
    from . import ops

    sig = [("precision", ops.PRECISION, ops.BACKWARD_PRECISION, ops.DECODER_PRECISION)]
    for k in sorted(batch):
        if k.startswith("_"):
            continue
        v = batch[k]
        if torch.is_tensor(v) and v.dim() > 0:
            sig.append((k, tuple(v.shape), str(v.dtype)))
        elif torch.is_tensor(v):
            sig.append((k, int(v)))
    return tuple(sig)


class GraphedSynthesis:
    def __init__(self, model, max_graphs: int = 64):
        self.model = model
        self.max_graphs = max_graphs
        self._cache: dict = {}

    def _capture(self, batch):
        model = self.model
        static_in = {k: (v.clone() if torch.is_tensor(v) and v.dim() > 0 and v.is_cuda else v) for k, v in batch.items()
                     if not k.startswith("_")}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):  # warm-up: allocator, cudaFuncSetAttribute, weight re-packs
                model(static_in, inference=True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        gc_was_enabled = gc.isenabled()
        gc.collect()
        gc.disable()  # destructors of collected CUDA objects must not run inside the capture (see GraphedTrainStep)
        try:
            with torch.cuda.graph(graph), torch.no_grad():
                static_out = model(static_in, inference=True)
        finally:
            if gc_was_enabled:
                gc.enable()
        return graph, static_in, static_out

    def __call__(self, batch, non_blocking: bool = True):
        """batch: dict with CUDA or pinned-host tensors.  Returns the (static) output dict of the replay —
        copy what must outlive the next call."""
        key = _signature(batch)
        entry = self._cache.get(key)
        if entry is None:
            if len(self._cache) >= self.max_graphs:
                self._cache.pop(next(iter(self._cache)))
            dev = next(self.model.parameters()).device
            dev_batch = {k: (v.to(dev) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in batch.items() if not k.startswith("_")}
            entry = self._capture(dev_batch)
            self._cache[key] = entry
        graph, static_in, static_out = entry
        for k, v in batch.items():
            if torch.is_tensor(v) and v.dim() > 0 and k in static_in:
                static_in[k].copy_(v, non_blocking=non_blocking)
        graph.replay()
        return static_out



class GraphedTrainStep:
    """One optimisation step — zero_grad, forward (aligner + MAS + encoder + variance adaptor + decoder + PostNet),
    the seven losses, backward, clip and AdamW — as ONE CUDA graph per batch shape (data-parallel runs: three graphs
    around two NCCL all-reduces on a communication stream, see `_exchange`): ≈ 850 kernel launches become one
    `cudaGraphLaunch`, so a step costs what the GPU needs, not what Python can enqueue.

    Everything that changes between steps lives in device memory: the batch (copied into the captured input
    buffers, straight from pinned host memory if that is where it is), the learning rate / Adam bias corrections
    and the dropout seed base (`FusedAdamW.begin_graph_step`).  The first `capture_after` (2) times a shape is seen the step runs
    eagerly (that is also the warm-up a capture needs); the next time it is captured and replayed.  Shapes are
    exact — padding is live in this model (BatchNorm statistics and the full-rectangle losses see it), so batches
    are never re-padded to share a graph.  The loss weights that depend on the epoch are part of the key.
    """

    def __init__(self, model, optimizer, scheduler=None, max_graphs: int = 16, overlap_wgrad: bool = True, capture_after: int = 2):
        self.model, self.opt, self.sched = model, optimizer, scheduler
        self.max_graphs = max_graphs
        self._cache: dict = {}
        self._seen: dict = {}           # batch shape → eager sights so far
        # A capture costs ≈ 0.4 s of host time (≈ 850 graph nodes to record and instantiate) — 15–35 eager steps' worth — so a
        # shape is captured only after `capture_after` eager sights: in a shape-diverse stream (measured: 332 distinct shapes in
        # 384 batches) capturing at the second sight made the stream 2.7× SLOWER than never capturing.
        self.capture_after = max(1, int(capture_after))
        # … and capturing backs off altogether when captured shapes are not coming back: a capture (with the allocator flush
        # torch.cuda.graph does around it) costs ≈ 1 s end to end, so after the first 4 captures a new shape is only captured
        # while the stream has shown at least 4 replays per capture (measured on 604 distinct shapes in 768 batches:
        # 90 ms per step with unconditional capturing against 24 ms eager).
        self._n_captures = 0
        self._n_replays = 0
        self._eager_only: set = set()
        self._pool = None
        self.overlap_wgrad = overlap_wgrad
        self.overlap_allreduce = True   # data-parallel runs: cut the backward at the decoder's input (see _early_params)
        self._sink = None
        self._early = None
        self._comm = None
        self.skip_exchange = False
        # every dropout kernel adds *seed_base to its by-value seed (fresh masks on graph replays).  The pointer is
        # process-wide inside the library: it is re-armed before every step of THIS runner (another runner may have
        # pointed it at its own counter) and cleared when the optimizer — the owner of the counter — goes away.
        self._token = next(_owner_tokens)
        self._arm_seed_base()
        weakref.finalize(optimizer, _clear_seed_base, optimizer.seed_base.data_ptr(), self._token)

    def _arm_seed_base(self) -> None:
        global _ARMED_SEED_BASE, _ARMED_OWNER
        ptr = self.opt.seed_base.data_ptr()
        if _ARMED_SEED_BASE != ptr or _ARMED_OWNER != self._token:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the dropout seed base must be armed before the capture starts")
            check(lib().fs2k_set_dropout_seed_base(ptr), "fs2k_set_dropout_seed_base")
            _ARMED_SEED_BASE, _ARMED_OWNER = ptr, self._token

    def _early_params(self):
        """Data-parallel runs: the parameters behind the decoder's input (decoder, mel_linear, PostNet).  Their gradients are
        final when the backward reaches that tensor, so their slice of the flat gradient is all-reduced while the variance
        adaptor, the aligner and the encoder are still back-propagating.  None when the backward is not cut."""
        if self._early is None:
            model, opt = self.model, self.opt
            early = []
            if opt.world_size() > 1 and self.overlap_allreduce:
                mods = [model.decoder, model.mel_linear] + ([model.postnet] if getattr(model, "postnet", None) is not None else [])
                early = [p for m in mods for p in m.parameters() if p.requires_grad]
                try:
                    opt.set_early_bucket(early)
                except ValueError:
                    early = []
            ids = {id(p) for p in early}
            self._early = (early, [p for p in opt._params if id(p) not in ids])
        return self._early

    def _forward_backward(self, batch, cut: bool = False):
        """zero_grad, forward, the seven losses and the backward — all of it, or (cut=True) only down to the decoder's
        input: then `pending` carries what `_backward_rest` needs.  Returns (detached losses, pending)."""
        from . import autograd_fns as fns

        model, opt = self.model, self.opt
        opt.zero_grad()
        early, _ = self._early_params() if cut else ([], None)
        holder = {}
        if early:
            def _cut(t):
                holder["up"] = t
                holder["leaf"] = t.detach().requires_grad_(True)
                return holder["leaf"]
            model._backward_cut = _cut
        prev = None
        if self.overlap_wgrad:
            # second stream: the forward-sum loss (started right after the aligner) and, in the backward, the weight /
            # bias gradients of every contraction, accumulated straight into the flat gradient
            if self._sink is None:
                self._sink = fns.WgradSink()
            prev = fns.set_wgrad_sink(self._sink)
        pending = None
        try:
            out = model(batch)
            losses = model.loss(out, batch, model.current_epoch)
            if self.overlap_wgrad:
                self._sink.fence()  # transposed weights / CTC loss prepared on the side stream during the forward
            if "leaf" in holder:
                losses["total"].backward(inputs=early + [holder["leaf"]], retain_graph=True)
                pending = (holder["up"], holder["leaf"], losses["total"])
            else:
                losses["total"].backward()
        finally:
            model._backward_cut = None
            if self.overlap_wgrad:
                fns.set_wgrad_sink(prev)
                self._sink.join()
        # detached: a loss that kept its autograd graph alive would also keep this step's AccumulateGrad nodes (and
        # their stream) alive into the next capture
        return {k: v.detach() for k, v in losses.items()}, pending

    def _backward_rest(self, pending) -> None:
        """Second phase of a cut backward: from the decoder's input (with the gradient phase one left on the cut) and from
        the losses that do not pass through the decoder, into every parameter outside the early bucket."""
        from . import autograd_fns as fns

        up, leaf, total = pending
        _, late = self._early_params()
        prev = fns.set_wgrad_sink(self._sink) if self.overlap_wgrad else None
        try:
            torch.autograd.backward([up, total], [leaf.grad, None], inputs=late)
        finally:
            if self.overlap_wgrad:
                fns.set_wgrad_sink(prev)
                self._sink.join()

    def _exchange(self, first: bool) -> None:
        """The gradient all-reduce of one bucket, on the communication stream (ordered after what the main stream has been
        given so far).  first=True: the early bucket; False: everything else."""
        opt = self.opt
        lo, hi = opt.early_bucket
        if self._comm is None:
            self._comm = torch.cuda.Stream()
        self._comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._comm):
            if first:
                opt.allreduce_range(lo, hi)
            else:
                opt.allreduce_range(0, lo)
                opt.allreduce_range(hi, opt.flat_g.numel())

    def _step_body(self, batch):
        """The whole step with eager launches.  Data-parallel runs issue exactly the collectives of the replayed step (same
        buckets, same order): ranks see their batch shapes in different orders, so one rank may be here — first sight of a
        shape — while another replays its graphs, and NCCL needs the same sequence of calls on every rank."""
        opt = self.opt
        if opt.world_size() > 1 and self._early_params()[0]:
            losses, pending = self._forward_backward(batch, cut=True)
            if pending is None:
                raise RuntimeError("the forward did not pass the decoder-input cut")
            self._exchange(first=True)
            self._backward_rest(pending)
            self._exchange(first=False)
            torch.cuda.current_stream().wait_stream(self._comm)
            opt.step(allreduce=False)
        else:
            losses, _ = self._forward_backward(batch)
            opt.step()
        return losses

    def _capture_pays(self) -> bool:
        return self._n_captures < 4 or self._n_replays >= 4 * self._n_captures

    def _key(self, batch):
        t = self.model.config.training
        return _signature(batch), min(self.model.current_epoch, t.attn_bin_loss_warmup_epochs)

    def __call__(self, batch, non_blocking: bool = True):
        """batch: collated dict (CUDA or pinned-host tensors).  Returns the dict of loss tensors (device scalars;
        static buffers of the graph — read or copy them before the next call)."""
        model, opt = self.model, self.opt
        self._arm_seed_base()
        key = self._key(batch)
        entry = self._cache.get(key)
        dev = opt.flat_p.device
        if entry is None and (self._seen.get(key, 0) < self.capture_after or key in self._eager_only or not self._capture_pays()):
            # first sights of this shape: plain eager step (validates the data, warms every lazy init)
            if len(self._seen) > 4096:
                self._seen.clear()
            self._seen[key] = self._seen.get(key, 0) + 1
            dev_batch = {k: (v.to(dev, non_blocking=non_blocking) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in batch.items()
                         if not k.startswith("_")}
            losses = self._step_body(dev_batch)
            if self.sched is not None:
                self.sched.step()
            return losses
        if entry is None:
            if len(self._cache) >= self.max_graphs:
                self._cache.pop(next(iter(self._cache)))
            static_in = {k: (v.to(dev).clone() if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in batch.items()
                         if not k.startswith("_")}
            va = model.variance_adaptor
            prev_validate, va.validate_durations = va.validate_durations, False  # the eager first sight validated
            opt.device_state = True
            torch.cuda.synchronize()
            # No cyclic garbage collection inside the capture: collecting an old model / runner there runs CUDA calls from its
            # destructors and finalizers (graph and pool destruction, the seed-base reset), which invalidates a capture in
            # progress ("operation failed due to a previous error during capture", seen intermittently between tests).
            gc_was_enabled = gc.isenabled()
            gc.collect()
            gc.disable()
            # data-parallel: the NCCL all-reduces stay OUTSIDE the graphs.  forward + backward down to the decoder's input
            # → [all-reduce of the decoder / PostNet gradients ‖ rest of the backward] → all-reduce of the rest → update;
            # single GPU: one graph for everything
            split = opt.world_size() > 1
            cut = split and bool(self._early_params()[0])
            graph = torch.cuda.CUDAGraph()
            graph_rest = torch.cuda.CUDAGraph() if cut else None
            graph2 = torch.cuda.CUDAGraph() if split else None
            try:
                with torch.cuda.graph(graph, pool=self._pool):
                    static_losses, pending = self._forward_backward(static_in, cut=cut)
                    if not split:
                        opt.step()
                if cut:
                    if pending is None:
                        raise RuntimeError("the forward did not pass the decoder-input cut")
                    with torch.cuda.graph(graph_rest, pool=graph.pool()):
                        self._backward_rest(pending)
                    del pending
                if split:
                    with torch.cuda.graph(graph2, pool=graph.pool()):
                        opt.step(allreduce=False)
            except Exception as e:  # a library op on the path that cannot be captured (e.g. cuDNN RNN autograd of the GST encoder)
                import warnings

                warnings.warn(f"GraphedTrainStep: this step cannot be captured into a CUDA graph ({type(e).__name__}: {e}); "
                              "running it with eager launches instead")
                self._eager_only.add(key)
                torch.cuda.synchronize()
                return self(batch, non_blocking)
            finally:
                if gc_was_enabled:
                    gc.enable()
                opt.device_state = False
                va.validate_durations = prev_validate
            if self._pool is None:
                self._pool = graph.pool()
            entry = (graph, static_in, static_losses, graph2, graph_rest)
            self._cache[key] = entry
            self._n_captures += 1
        else:
            self._n_replays += 1
        graph, static_in, static_losses, graph2, graph_rest = entry
        for k, v in batch.items():
            if torch.is_tensor(v) and v.dim() > 0 and k in static_in:
                static_in[k].copy_(v, non_blocking=non_blocking)
        opt.begin_graph_step()
        graph.replay()
        if self.skip_exchange and graph2 is not None:   # measurement only (bench.py): the same graphs without the all-reduces
            if graph_rest is not None:
                graph_rest.replay()
            graph2.replay()
        elif graph_rest is not None:
            self._exchange(first=True)      # decoder / PostNet gradients travel ...
            graph_rest.replay()             # ... while the variance adaptor, aligner and encoder back-propagate
            self._exchange(first=False)
            torch.cuda.current_stream().wait_stream(self._comm)
            graph2.replay()
        elif graph2 is not None:
            opt.allreduce_grads()
            graph2.replay()
        if self.sched is not None:
            self.sched.step()
        return static_losses
