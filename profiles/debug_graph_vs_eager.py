"""Debug aid: loss trajectories of the bf16 training step, eager vs CUDA-graph replay, with individual kernels switched back
to their predecessors (MAS wavefront, row-panel GEMM, tensor-core attention) to find which one makes the two paths differ."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import torch

from helpers import case_batch, load_case
from test_model_gpu import build_model
from fastspeech2_lightning_b200 import ops, autograd as ag
from fastspeech2_lightning_b200._lib import lib

DEV = "cuda:0"
meta, _ = load_case("train_bn")
batch = case_batch(meta, DEV)


def run(mode, steps=4):
    torch.manual_seed(0)
    model = build_model(meta)
    model.fused_grad_clip = 1.0
    model.configure_optimizers()
    out = []
    for i in range(steps):
        losses = model.optimization_step(batch, use_cuda_graph=(mode == "graph"))
        out.append({k: round(float(v), 5) for k, v in losses.items()})
    return out


def show(tag):
    a, b, c = run("eager"), run("eager"), run("graph")
    print(f"--- {tag}")
    for i, (x, y, z) in enumerate(zip(a, b, c)):
        bad_ee = [k for k in x if abs(x[k] - y[k]) > 2e-3 * max(1, abs(x[k]))]
        bad_eg = [k for k in x if abs(x[k] - z[k]) > 2e-3 * max(1, abs(x[k]))]
        print(f"step {i}: eager/eager differ {bad_ee} eager/graph differ {bad_eg}")
        if bad_ee or bad_eg:
            print("   eager ", x)
            print("   eager2", y)
            print("   graph ", z)



ops.set_precision("bf16")
variant = os.environ.get("VARIANT", "all")
print("VARIANT", variant, "FS2K_PDL", os.environ.get("FS2K_PDL", "1"))
if variant == "nosink":
    from fastspeech2_lightning_b200 import graphs
    orig_init = graphs.GraphedTrainStep.__init__
    def init(self, *a, **k):
        orig_init(self, *a, **k)
        self.overlap_wgrad = False
    graphs.GraphedTrainStep.__init__ = init
if variant == "fwdonly_tc":   # tensor-core forward, fp32 SIMT backward is impossible (different saved tensors): skip
    pass
show(f"bf16 [{variant}]")
