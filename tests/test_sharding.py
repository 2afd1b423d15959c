"""World-size-2 ``gloo`` test of the multi-rank path on CPU: contiguous sharding + one gather."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))


def test_shard_ranges_cover_and_preserve_order():
    from selene_b200.sharding import shard_range, shard_sizes
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(shard_sizes(n, world)) - min(shard_sizes(n, world)) <= 1


def _worker(rank, world, port, n_total, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from selene_b200.sharding import shard_range, gather_tiles
    lo, hi = shard_range(n_total, rank, world)
    full = torch.arange(n_total * 5, dtype=torch.float32).reshape(n_total, 5)
    local = full[lo:hi].clone() * 1.0           # this rank's "score tile"
    everyone = gather_tiles(local, n_total)
    assert torch.equal(everyone, full)
    root = gather_tiles(local, n_total, dst=0)
    if rank == 0:
        assert torch.equal(root, full)
        np.save(out_path, root.numpy())
    else:
        assert root is None
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 11])
def test_gather_tiles_gloo_world2(tmp_path, n_total):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + n_total
    out = str(tmp_path / "root.npy")
    mp.spawn(_worker, args=(2, port, n_total, out), nprocs=2, join=True)
    got = np.load(out)
    assert got.shape == (n_total, 5) and got[-1, -1] == n_total * 5 - 1
