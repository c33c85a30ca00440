"""ORACLE-side test infrastructure (not product code): run the UNMODIFIED reference `fs2.model.FastSpeech2` on the
host CPU — the `kind: "reference"` arm of bench.py (`--impl reference`, `cpu_baseline`).

The module is built through `oracle/ref_shim.py` (stubs for everyvoice / pytorch_lightning / matplotlib only; every
numerical line is the reference's own: fs2/model.py, variance_adaptor.py, layers.py, attn/*, loss.py, noam.py,
torchaudio's Conformer, numba MAS) from `/root/reference` or the build-time copy under `oracle/_ref/`.  Training runs
with the reference's dropout ACTIVE (`model.train()`: Conformer dropout, predictor dropout, PostNet's F.dropout 0.5),
`clip_grad_norm_(1.0)` (fs2/cli/train.py:38), torch.optim.AdamW + NoamLR (fs2/model.py:530-549).
"""
from __future__ import annotations

import os
import time

import torch

from . import ref_shim


def available() -> bool:
    return ref_shim.reference_available()


class ReferenceRunner:
    def __init__(self, config, state_dict=None, lang2id=None, speaker2id=None, stats=None, threads=None, device="cpu"):
        from fastspeech2_lightning_b200 import synthetic

        self.ref = ref_shim.reference_modules()
        self.threads = threads or (os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        torch.manual_seed(1234)
        self.model = self.ref.model.FastSpeech2(config, stats=stats or synthetic.DEFAULT_STATS, lang2id=lang2id or {}, speaker2id=speaker2id or {})
        if state_dict is not None:
            self.model.load_state_dict({k: v.detach().cpu() for k, v in state_dict.items()}, strict=True)
        self.device = torch.device(device)
        self.model.to(self.device)  # "cpu": the reference's CPU path; a CUDA device: stock PyTorch eager on that GPU
        self.control = self.ref.model.InferenceControl()
        self.optimizer = self.scheduler = None

    # ---- synthesis forward (eval, no grad): fs2/model.py:153-268 ----
    def synthesize(self, batch):
        self.model.eval()
        with torch.no_grad():
            return self.model(batch, control=self.control, inference=True)

    # ---- learned-alignment forward (eval, no grad; aligner + MAS run: BASELINE configs[4]) ----
    def align_forward(self, batch):
        self.model.eval()
        with torch.no_grad():
            return self.model(batch, control=self.control)

    # ---- one optimisation step: fs2/model.py:384-390 + Lightning's clip + configure_optimizers ----
    def train_step(self, batch, clip: float = 1.0):
        m = self.model
        m.train()
        if self.optimizer is None:
            cfg = m.configure_optimizers()
            self.optimizer, self.scheduler = cfg[0][0], cfg[1][0]["scheduler"]
        self.optimizer.zero_grad()
        out = m(batch, control=self.control)
        losses = m.loss(out, batch, m.current_epoch)
        losses["total"].backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), clip)
        self.optimizer.step()
        self.scheduler.step()
        return float(losses["total"])

    def time_steps(self, fn, batches, steps: int, warmup: int):
        for i in range(warmup):
            fn(batches[i % len(batches)])
        t0 = time.perf_counter()
        for i in range(steps):
            fn(batches[i % len(batches)])
        return time.perf_counter() - t0
