"""Data-parallel training on real GPUs (needs ≥ 2 devices; skipped on a single-GPU box — `gpurun --gpus 2` runs it):
a 2-rank step through the path the bench uses (`FastSpeech2.optimization_step`: graph replay(s) around the NCCL all-reduce of
the flat gradient, 1/world folded into the clip + AdamW launch) must equal the 1-rank step on the concatenated batch.

For the equality to be exact in exact arithmetic the two half-batches are padded to the same (T, F) (the losses are means
over the padded tensors, loss.py:19-126), BatchNorm uses its running statistics (Lightning DDP keeps batch statistics per
replica — with them the two runs differ by construction) and the epoch-0 weight of the attention binarisation loss is 0
(it is a ratio of sums, not a mean).
"""
import datetime
import os
import sys
import tempfile
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _model_and_batches(device, precision):
    sys.path.insert(0, str(ROOT))
    from fastspeech2_lightning_b200 import ops, synthetic
    from fastspeech2_lightning_b200.fs2.config import FastSpeech2Config
    from fastspeech2_lightning_b200.fs2.model import FastSpeech2

    ops.set_precision(precision)
    nodrop = dict(encoder=dict(dropout=0.0), decoder=dict(dropout=0.0),
                  variance_predictors=dict(energy=dict(dropout=0.0), pitch=dict(dropout=0.0), duration=dict(dropout=0.0)))
    cfg = FastSpeech2Config(model=nodrop)
    model = FastSpeech2(cfg, stats=synthetic.DEFAULT_STATS)
    synthetic.fill_weights_(model, seed=77)
    model.postnet.dropout_in_training = False
    model.eval()                     # BatchNorm running statistics, no dropout: see the module docstring
    model.current_epoch = 0
    model = model.to(device)
    model.variance_adaptor.validate_durations = False
    model.fused_grad_clip = 1.0
    halves = [synthetic.make_batch(2, (16, 16), seed=900 + r, learn_alignment=True, dur_range=(5, 5), src_lens=[16, 9 + 3 * r]) for r in range(2)]
    return model, halves


def _cat(halves):
    out = {}
    for k, v in halves[0].items():
        w = halves[1][k]
        if torch.is_tensor(v) and v.dim() > 0:
            out[k] = torch.cat([v, w], dim=0)
        elif isinstance(v, list):
            out[k] = v + w
        else:
            out[k] = v
    return out


def _worker(rank, world, port, precision, ref_path, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    torch.distributed.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
    model, halves = _model_and_batches(dev, precision)
    model.configure_optimizers()
    from fastspeech2_lightning_b200 import synthetic

    batch = synthetic.batch_to(halves[rank], dev)
    losses = []
    for _ in range(4):               # two eager sights, capture + replay, replay
        losses.append(float(model.optimization_step(batch)["total"]))
    flat = model.optimizer.flat_p.detach().cpu()
    if rank == 0:
        ref = torch.load(ref_path)
        err = float((flat - ref["flat"]).abs().max() / ref["flat"].abs().max())
        torch.save({"err": err, "losses": losses, "ref_losses": ref["losses"]}, out_path)
    torch.cuda.synchronize()
    torch.distributed.barrier()
    os._exit(0)


@pytest.mark.parametrize("precision,tol", [("tf32x3", 2e-5), ("bf16", 2e-3)])
def test_two_rank_step_equals_one_rank_step_on_the_concatenated_batch(precision, tol):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp

    from fastspeech2_lightning_b200 import ops, synthetic

    dev = torch.device("cuda", 0)
    model, halves = _model_and_batches(dev, precision)
    try:
        model.configure_optimizers()
        full = synthetic.batch_to(_cat(halves), dev)
        assert int(halves[0]["max_mel_len"]) == int(halves[1]["max_mel_len"]) == 80
        ref_losses = [float(model.optimization_step(full)["total"]) for _ in range(4)]
        with tempfile.TemporaryDirectory() as td:
            ref_path, out_path = os.path.join(td, "ref.pt"), os.path.join(td, "out.pt")
            torch.save({"flat": model.optimizer.flat_p.detach().cpu(), "losses": ref_losses}, ref_path)
            del model
            torch.cuda.empty_cache()
            port = 29500 + (os.getpid() % 2000)
            ctx = mp.spawn(_worker, args=(2, port, precision, ref_path, out_path), nprocs=2, join=True)
            res = torch.load(out_path)
    finally:
        ops.set_precision("tf32x3")
    print(f"{precision}: 2-rank vs 1-rank parameters after 4 steps: max err / max|p| = {res['err']:.2e}; "
          f"rank-0 losses {res['losses']} vs full-batch {res['ref_losses']}")
    assert res["err"] <= tol, res


def _worker_mixed(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    torch.distributed.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
    model, _ = _model_and_batches(dev, "bf16")
    model.configure_optimizers()
    from fastspeech2_lightning_b200 import synthetic

    def mk(n, seed):
        return synthetic.batch_to(synthetic.make_batch(2, (n, n), seed=seed, learn_alignment=True, dur_range=(5, 5), src_lens=[n, n - 5]), dev)

    # rank 0 alternates two shapes (eager, eager, capture, capture, replay, replay); rank 1 repeats one (eager, capture, replay …)
    seq = [mk(16, 1), mk(12, 2)] * 3 if rank == 0 else [mk(14, 3)] * 6
    for b in seq:
        loss = float(model.optimization_step(b)["total"])
        assert loss == loss
    flat = model.optimizer.flat_p.detach()
    both = [torch.empty_like(flat) for _ in range(world)]
    torch.distributed.all_gather(both, flat)
    if rank == 0:
        torch.save({"diff": float((both[0] - both[1]).abs().max()), "cache": len(model._train_runner._cache)}, out_path)
    torch.cuda.synchronize()
    torch.distributed.barrier()
    os._exit(0)


def test_ranks_in_different_step_modes_issue_the_same_collectives():
    """Ranks meet their batch shapes in different orders, so at the same step one rank may run its eager first sight while the
    other replays captured graphs: both must issue the same NCCL calls (same buckets, same order) — a mismatch hangs the job
    (seen at 8 GPUs with one rank drawing the same shape twice).  The replicas stay bit-identical."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp

    with tempfile.TemporaryDirectory() as td:
        out_path = os.path.join(td, "out.pt")
        mp.spawn(_worker_mixed, args=(2, 29700 + (os.getpid() % 2000), out_path), nprocs=2, join=True)
        res = torch.load(out_path)
    assert res["diff"] == 0.0 and res["cache"] == 2, res
