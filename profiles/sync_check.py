"""Lists every host synchronisation inside one eager training step (torch's sync debug mode), so the step can be
captured into a CUDA graph.  Run on the GPU box: python profiles/sync_check.py"""
import sys
import warnings
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg, model = bench.build_train_model(dev)
model.configure_optimizers()
from fastspeech2_lightning_b200 import synthetic

wl = bench.WORKLOADS["train_c2"]
batch = synthetic.batch_to(bench.make_train_batches(wl, 1, 0)[0], dev)
model.optimization_step(batch, use_cuda_graph=False)
torch.cuda.synchronize()
model.variance_adaptor.validate_durations = False
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    model.optimization_step(batch, use_cuda_graph=False)
torch.cuda.set_sync_debug_mode("default")
print("synchronising calls in one step:", len(w))
for x in w:
    print(" ", x.filename, x.lineno, str(x.message)[:100])
