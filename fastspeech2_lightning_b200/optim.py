"""Flat-buffer AdamW with fused gradient clipping (and the data-parallel gradient all-reduce).

All trainable parameters are re-pointed into ONE contiguous fp32 buffer and their `.grad`s into a second
one, so that a training-step tail is: one NCCL all-reduce of the flat gradient (multi-GPU), one Σg²
reduction, one fused clip + AdamW launch — instead of 400+ per-tensor launches.  Same update rule and
hyper-parameters as the reference's `torch.optim.AdamW` (fs2/model.py:530-537) with Lightning's
`gradient_clip_val` (fs2/cli/train.py:38); it is a `torch.optim.Optimizer`, so `NoamLR` and Lightning drive it.
"""
from __future__ import annotations

import math
import struct
import weakref

import torch

from . import ops
from ._lib import check, lib


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None,
                 process_group=None):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("no trainable parameters")
        if any(not p.is_cuda or p.dtype != torch.float32 for p in params):
            raise ValueError("FusedAdamW needs fp32 CUDA parameters (no CPU path)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = max_grad_norm
        self.process_group = process_group
        dev = params[0].device
        n = sum(p.numel() for p in params)
        # slices aligned to 8 elements: 16 bytes in the bf16 shadow too (TMA base addresses), 32 bytes in fp32
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 7) // 8 * 8
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        # bf16 shadow of the parameters (bf16 arithmetic mode): the update kernel writes it next to the fp32 masters,
        # so the GEMMs' weight operands need no cast pass; created on first use (ops.bf16_weight)
        self.flat_p16 = None
        self._p16_sig = -1  # Σ parameter version counters at the last sync
        self.early_bucket = None   # (lo, hi) of the gradients that are final first in the backward (set_early_bucket)
        self.comm_bf16 = True      # bf16 mode: all-reduce the gradient as bf16
        self._g16 = None
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self.flat_p[o: o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_g[o: o + p.numel()].view_as(p)
                p._fs2k_flat_owner = weakref.ref(self)
        self._params, self._offs, self.numel = params, offs, n
        self._step = 0
        # device copies of the per-step scalars {lr, 1−β1^t, sqrt(1−β2^t)} and the dropout seed base: a captured
        # (CUDA-graph) step reads them instead of by-value arguments — see begin_graph_step()
        self.step_state = torch.zeros(4, dtype=torch.float32, device=dev)
        self.seed_base = torch.zeros(1, dtype=torch.int64, device=dev)
        self.device_state = False  # True while a step is being captured: step() must not touch host scalars

    def zero_grad(self, set_to_none: bool = False):
        """Gradients stay views of the flat buffer (never set to None): one memset."""
        self.flat_g.zero_()
        for p, o in zip(self._params, self._offs):
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                p.grad = self.flat_g[o: o + p.numel()].view_as(p)

    def grad_norm(self) -> torch.Tensor:
        stream = torch.cuda.current_stream().cuda_stream
        check(lib().fs2k_sumsq(self.flat_g.data_ptr(), self.flat_g.numel(), self._sumsq.data_ptr(), stream), "fs2k_sumsq")
        ops._count()
        return self._sumsq.sqrt()

    def _bump_versions(self) -> None:
        """The update kernel writes the parameters through raw pointers, which torch's version counters do not see:
        bump them, so that anything cached per parameter version (the re-packed conv weights in ops) is refreshed."""
        torch.autograd.graph.increment_version(self._params)

    def world_size(self) -> int:
        if self.process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return torch.distributed.get_world_size(self.process_group)
        return 1

    @torch.no_grad()
    def allreduce_grads(self) -> None:
        """The data-parallel exchange: one (sum) all-reduce of the flat gradient; the 1/world mean folds into the update."""
        if self.world_size() > 1:
            self.allreduce_range(0, self.flat_g.numel())

    def set_early_bucket(self, params) -> None:
        """Mark the parameters whose gradients are final first in the backward (decoder, mel_linear, PostNet): a contiguous
        slice [lo, hi) of the flat gradient that can be all-reduced while the rest of the backward still runs."""
        ids = {id(p) for p in params}
        idx = [i for i, p in enumerate(self._params) if id(p) in ids]
        if not idx or idx != list(range(idx[0], idx[-1] + 1)):
            raise ValueError("the early bucket must be a contiguous run of the optimizer's parameters")
        lo = self._offs[idx[0]]
        hi = self._offs[idx[-1] + 1] if idx[-1] + 1 < len(self._offs) else self.flat_g.numel()
        self.early_bucket = (lo, hi)

    @torch.no_grad()
    def allreduce_range(self, lo: int, hi: int) -> None:
        """Sum all-reduce of flat_g[lo:hi] on the current stream.  bf16 arithmetic mode: the gradient travels as bf16 (half the
        bytes over NVLink; the fp32 master gradient is rounded once on the way out and restored on the way back)."""
        if hi <= lo or self.world_size() == 1:
            return
        g = self.flat_g[lo:hi]
        if ops.PRECISION == "bf16" and self.comm_bf16:
            if self._g16 is None:
                self._g16 = torch.empty(self.flat_g.numel(), dtype=torch.bfloat16, device=self.flat_g.device)
            g16 = self._g16[lo:hi]
            stream = torch.cuda.current_stream().cuda_stream
            check(lib().fs2k_cast_bf16(g.data_ptr(), g.numel(), g16.data_ptr(), None, stream), "fs2k_cast_bf16")
            torch.distributed.all_reduce(g16, group=self.process_group)
            check(lib().fs2k_cast_f32(g16.data_ptr(), g.numel(), g.data_ptr(), stream), "fs2k_cast_f32")
            ops._count(2)
        else:
            torch.distributed.all_reduce(g, group=self.process_group)

    @torch.no_grad()
    def step(self, closure=None, allreduce: bool = True):
        loss = closure() if closure is not None else None
        world = self.world_size()
        if world > 1 and allreduce:
            self.allreduce_grads()
        g = self.param_groups[0]
        stream = torch.cuda.current_stream().cuda_stream
        sumsq = None
        if self.device_state:
            if self.max_grad_norm is not None:
                check(lib().fs2k_sumsq(self.flat_g.data_ptr(), self.flat_g.numel(), self._sumsq.data_ptr(), stream), "fs2k_sumsq")
                ops._count()
                sumsq = self._sumsq.data_ptr()
            check(lib().fs2k_adamw_step_dev(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(), self.flat_v.data_ptr(),
                                            self.flat_p.numel(), self.step_state.data_ptr(), float(g["betas"][0]), float(g["betas"][1]),
                                            float(g["eps"]), float(g["weight_decay"]), float(self.max_grad_norm or 0.0), 1.0 / world,
                                            sumsq, self._p16_ptr(), stream), "fs2k_adamw_step_dev")
            ops._count()
            return loss
        self._adopt_foreign_grads()
        self._step += 1
        self._bump_versions()
        if self.max_grad_norm is not None:
            check(lib().fs2k_sumsq(self.flat_g.data_ptr(), self.flat_g.numel(), self._sumsq.data_ptr(), stream), "fs2k_sumsq")
            ops._count()
            sumsq = self._sumsq.data_ptr()
        check(lib().fs2k_adamw_step(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(), self.flat_v.data_ptr(),
                                    self.flat_p.numel(), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                    float(g["weight_decay"]), self._step, float(self.max_grad_norm or 0.0), 1.0 / world, sumsq,
                                    self._p16_ptr(), stream),
              "fs2k_adamw_step")
        ops._count()
        self._mark_p16_synced()
        return loss

    def _adopt_foreign_grads(self) -> None:
        """The update reads the flat gradient buffer.  Anything that replaced a parameter's `.grad` (nn.Module.zero_grad /
        a zero_grad hook with set_to_none, DDP's gradient_as_bucket_view, a hand-assigned tensor) made autograd accumulate
        somewhere else: copy such gradients into their slice and re-point `.grad` at it.  A parameter whose grad is None
        gets a zero slice — unlike torch.optim.AdamW, which skips it entirely, it still receives weight decay and moment
        decay (every parameter of this model receives a gradient in every step, so the two only differ for frozen-by-hand
        parameters, which should be excluded from the optimizer instead)."""
        base = self.flat_g.data_ptr()
        for p, o in zip(self._params, self._offs):
            g = p.grad
            if g is not None and g.data_ptr() == base + 4 * o:
                continue
            view = self.flat_g[o: o + p.numel()].view_as(p)
            if g is None:
                view.zero_()
            else:
                view.copy_(g)
            p.grad = view

    # ---- checkpoint format: torch.optim.AdamW's (per-parameter step / exp_avg / exp_avg_sq) -----------------------------------
    def state_dict(self):
        """Same layout as the reference's torch.optim.AdamW (fs2/model.py:530-537), sliced out of the flat moment buffers,
        so Lightning checkpoints (`optimizer_states`) are interchangeable in both directions."""
        sd = super().state_dict()
        state = {}
        if self._step > 0:
            for i, (p, o) in enumerate(zip(self._params, self._offs)):
                n = p.numel()
                state[i] = {"step": torch.tensor(float(self._step)),
                            "exp_avg": self.flat_m[o: o + n].view_as(p).clone(),
                            "exp_avg_sq": self.flat_v[o: o + n].view_as(p).clone()}
        sd["state"] = state
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        state = state_dict.get("state", {})
        super().load_state_dict({"state": {}, "param_groups": state_dict["param_groups"]})
        self.flat_m.zero_()
        self.flat_v.zero_()
        steps = set()
        for i, (p, o) in enumerate(zip(self._params, self._offs)):
            st = state.get(i, state.get(str(i)))
            if st is None:
                continue
            n = p.numel()
            self.flat_m[o: o + n].view_as(p).copy_(st["exp_avg"])
            self.flat_v[o: o + n].view_as(p).copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"FusedAdamW keeps one step count for all parameters; the checkpoint has {sorted(steps)}")
        self._step = steps.pop() if steps else 0

    # ---- bf16 shadow of the weights ----------------------------------------------------------------------------------
    def _p16_ptr(self):
        return None if self.flat_p16 is None else self.flat_p16.data_ptr()

    def _mark_p16_synced(self) -> None:
        """Called right after _bump_versions (every counter + 1) when the update kernel also wrote the shadow."""
        if self.flat_p16 is not None and self._p16_sig >= 0:
            self._p16_sig += len(self._params)

    def _version_sig(self) -> int:
        return sum(p._version for p in self._params)

    def bf16_shadow(self) -> torch.Tensor:
        """bf16 copy of the flat parameter buffer, kept current by the update kernel.  Anything else that writes the
        parameters (load_state_dict, an in-place fill) bumps their version counters: one cast launch re-syncs.  Called once
        per model forward (ops.refresh_bf16_shadows), not per weight."""
        if self.flat_p16 is None:
            self.flat_p16 = torch.empty(self.flat_p.numel(), dtype=torch.bfloat16, device=self.flat_p.device)
            ops.register_bf16_shadow(self)
        if not torch.cuda.is_current_stream_capturing():
            sig = self._version_sig()
            if sig != self._p16_sig:
                check(lib().fs2k_cast_bf16(self.flat_p.data_ptr(), self.flat_p.numel(), self.flat_p16.data_ptr(), None,
                                           torch.cuda.current_stream().cuda_stream), "fs2k_cast_bf16")
                ops._count()
                self._p16_sig = sig
        return self.flat_p16


    def begin_graph_step(self) -> None:
        """Host side of one replayed step: advance the step count and publish this step's learning rate, bias
        corrections and a fresh dropout seed base to the device (one tiny launch, arguments by value)."""
        g = self.param_groups[0]
        self._step += 1
        self._bump_versions()
        self._mark_p16_synced()  # the replayed update kernel writes the bf16 shadow too
        self._opt_called = True  # LRScheduler's "scheduler.step() before optimizer.step()" check
        # the betas as the kernels see them (fp32), the corrections in double: exactly what fs2k_adamw_step computes (optimizer.cu)
        b1, b2 = (struct.unpack("f", struct.pack("f", b))[0] for b in g["betas"])
        seed = int(torch.randint(0, 2**62, (1,)).item())
        check(lib().fs2k_set_step_state(self.step_state.data_ptr(), self.seed_base.data_ptr(), float(g["lr"]), 1.0 - math.pow(b1, self._step),
                                        math.sqrt(1.0 - math.pow(b2, self._step)), seed, torch.cuda.current_stream().cuda_stream),
              "fs2k_set_step_state")
        ops._count()
