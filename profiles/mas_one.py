"""A few launches of the MAS entry point at the long-utterance shape (BASELINE configs[4]: 8000 frames x 1000 phonemes) —
the command the `ncu --set full` capture wraps.  usage: python profiles/mas_one.py [F T B]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from fastspeech2_lightning_b200 import ops

F, T, B = (int(v) for v in (sys.argv[1:4] + ["8000", "1000", "1"][len(sys.argv) - 1:]))
g = torch.Generator(device="cuda").manual_seed(0)
soft = torch.softmax(torch.randn(B, 1, F, T, device="cuda", generator=g), -1)
il = torch.full((B,), T, dtype=torch.int32, device="cuda")
ol = torch.full((B,), F, dtype=torch.int32, device="cuda")
for _ in range(3):
    path, dur, hard = ops.mas(soft, il, ol, take_log=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.mas(soft, il, ol, take_log=True)
e1.record()
torch.cuda.synchronize()
print("ok", F, T, B, "ms per call", e0.elapsed_time(e1) / 10, "durations sum", int(dur.sum()))
