"""world_size-2 `gloo` tests (CPU) of the data-parallel host logic: utterance sharding, flat-gradient mean,
max-over-ranks timing, per-rank synthetic batches."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fastspeech2_lightning_b200 import parallel, synthetic

    # 1) flat gradient: every rank contributes rank+1 → mean = (1+2)/2
    flat = torch.full((1000,), float(rank + 1))
    parallel.allreduce_mean_(flat)
    assert torch.allclose(flat, torch.full((1000,), 1.5))
    # 2) timing = max over ranks, throughput numerator = sum over ranks
    assert parallel.max_over_ranks(10.0 + rank) == 11.0
    assert parallel.sum_over_ranks(100.0 * (rank + 1)) == 300.0
    # 3) utterance sharding: disjoint, covering, length-sorted batches
    g = torch.Generator().manual_seed(0)
    lengths = torch.randint(20, 201, (256,), generator=g).tolist()
    mine = parallel.shard_utterances(lengths, rank, world, 32)
    assert len(mine) == 4 and all(len(b) == 32 for b in mine)
    for b in mine:
        ls = [lengths[i] for i in b]
        assert ls == sorted(ls, reverse=True)
    flat_idx = torch.tensor(sorted(i for b in mine for i in b))
    gathered = [torch.zeros_like(flat_idx) for _ in range(world)]
    dist.all_gather(gathered, flat_idx)
    allidx = torch.cat(gathered).tolist()
    assert sorted(allidx) == list(range(256))
    # 4) ranks draw different synthetic batches from the same generator
    b = synthetic.make_batch(4, (10, 12), seed=1234 + 1000 * rank)
    sums = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(sums, b["text"].long().sum().reshape(1))
    assert sums[0].item() != sums[1].item()
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_data_parallel_host_logic_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_sharding_single_rank_is_identity_partition():
    from fastspeech2_lightning_b200.parallel import shard_utterances

    lengths = [5, 9, 1, 7, 7, 3]
    batches = shard_utterances(lengths, 0, 1, 4)
    assert sorted(i for b in batches for i in b) == list(range(6))
    assert batches[0] == [1, 3, 4, 0]
