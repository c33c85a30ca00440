"""Two eager synthesis steps of the C1 workload (first = warm-up) — the command the ncu captures wrap.
usage: python profiles/one_step.py [precision] ; prints the number of kernel launches per step."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import bench
from fastspeech2_lightning_b200 import ops, synthetic

ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
wl = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "synth_c1"]
dev = torch.device("cuda:0")
cfg, model = bench.build_model(wl, dev)
batch = synthetic.batch_to(bench.make_batches(wl, 1, 0)[0], dev)
with torch.no_grad():
    model(batch, inference=True)
    torch.cuda.synchronize()
    n0 = ops.launch_count
    model(batch, inference=True)
    torch.cuda.synchronize()
print("launches_per_step", ops.launch_count - n0, "first_step", n0)
