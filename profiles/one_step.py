"""Two eager steps of a bench workload (first = warm-up) — the command the ncu captures wrap.
usage: python profiles/one_step.py [precision] [workload]   (workload: synth_c1 | synth_c4 | train_c2)
prints the number of kernel launches per step."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import bench
from fastspeech2_lightning_b200 import ops, synthetic

ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
name = sys.argv[2] if len(sys.argv) > 2 else "synth_c1"
wl = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
if wl["kind"] == "train":
    cfg, model = bench.build_train_model(dev)
    (opt,), (sched,) = model.configure_optimizers()
    batch = synthetic.batch_to(bench.make_train_batches(wl, 1, 0)[0], dev)

    def step():
        model.optimization_step(batch, use_cuda_graph=False)  # eager launches, wgrad on the side stream
else:
    cfg, model = bench.build_model(wl, dev)
    batch = synthetic.batch_to(bench.make_batches(wl, 1, 0)[0], dev)

    def step():
        with torch.no_grad():
            model(batch, inference=True)

step()
torch.cuda.synchronize()
n0 = ops.launch_count
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("launches_per_step", ops.launch_count - n0, "first_step", n0)
