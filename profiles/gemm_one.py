"""One representative tensor-core GEMM (decoder shape: M = 32·500, K = N = 256, bias + residual), a few launches —
the command the `ncu --set full` capture of gemm_tc wraps."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from fastspeech2_lightning_b200 import ops

big = torch.randn(32, 500, 256, device="cuda")
w = torch.randn(256, 256, device="cuda") / 16
b = torch.randn(256, device="cuda")
y = big
for _ in range(6):
    y = ops.gemm(y, w, b, residual=big)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
