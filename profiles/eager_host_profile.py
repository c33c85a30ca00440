"""Host-side profile (cProfile) of eager training steps — where the Python time of a first-sight (uncaptured) step goes.
usage: python profiles/eager_host_profile.py [precision] [out.prof]"""
import cProfile
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import bench
from fastspeech2_lightning_b200 import ops, synthetic

ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "bf16")
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/eager_host.prof"
wl = bench.WORKLOADS["train_c2"]
dev = torch.device("cuda:0")
cfg, model = bench.build_train_model(dev)
model.configure_optimizers()
batches = [synthetic.batch_to(b, dev) for b in bench.make_train_batches(wl, 4, 0)]
for i in range(3):
    model.optimization_step(batches[i % 4], use_cuda_graph=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(4):
    model.optimization_step(batches[i % 4], use_cuda_graph=False)
torch.cuda.synchronize()
print("eager ms/step (wall)", (time.perf_counter() - t0) / 4 * 1e3)
pr = cProfile.Profile()
pr.enable()
for i in range(4):
    model.optimization_step(batches[i % 4], use_cuda_graph=False)
torch.cuda.synchronize()
pr.disable()
pr.dump_stats(out)
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
