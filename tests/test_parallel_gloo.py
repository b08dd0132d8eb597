"""The N > 1 path on the CPU: world size 2 over gloo (127.0.0.1), shard arithmetic and the metric all-reduce."""
import os
import socket

import pytest
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

from aps_b200.parallel import reduce_metrics, shard_bounds


def test_shard_bounds_partition_every_batch():
    for n in (0, 1, 7, 64, 255, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(13, rank, world)
    frames = float(sum(397 for _ in range(lo, hi)))
    red = reduce_metrics({"elapsed_ms": 10.0 + rank, "frames": frames, "utts": float(hi - lo)})
    if rank == 0:
        th.save(red, out)
    dist.barrier()
    dist.destroy_process_group()


def test_metric_reduction_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "red.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    red = th.load(out)
    assert red == {"elapsed_ms": 11.0, "frames": 13 * 397.0, "utts": 13.0}


def test_reduce_is_identity_without_group():
    assert reduce_metrics({"elapsed_ms": 3.0, "frames": 5.0}) == {"elapsed_ms": 3.0, "frames": 5.0}
