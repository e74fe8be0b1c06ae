"""World-size-2 gloo tests (CPU) of the batch-sharding host logic used for N > 1 GPUs."""
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


def _worker(rank, world, port, n_items, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from highres_net_b200 import distributed as hd
        g = torch.Generator().manual_seed(0)
        lrs = torch.rand(n_items, 3, 4, 4, generator=g)
        alphas = torch.ones(n_items, 3)

        def fake_model(x, a):                       # stand-in with the HRNet signature: (b, L, H, W) -> (b, 1, 3H, 3W)
            assert x.shape[0] > 0, "hrn_forward rejects an empty batch: ranks without work must not call the model"
            return (x * a[:, :, None, None]).sum(1, keepdim=True).repeat_interleave(3, 2).repeat_interleave(3, 3)

        l_lrs, l_alphas, (lo, hi) = hd.shard_batch(lrs, alphas)
        assert (lo, hi) == hd.shard_range(n_items, rank, world)
        sr, scores, xy = hd.sharded_forward_and_score(fake_model, lrs, alphas)
        assert scores is None and xy is None
        ref = fake_model(lrs, alphas)
        # a sub-group that holds only this rank: the shard must follow the GROUP's rank / size, not the global ones
        solo = [dist.new_group([r]) for r in range(world)][rank]
        sr_solo, _, _ = hd.sharded_forward_and_score(fake_model, lrs, alphas, group=solo)
        assert torch.equal(sr_solo, ref)
        results[rank] = bool(torch.equal(sr, ref)) and sr.shape[0] == n_items
        row = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1)
        gathered = hd.gather_batch(row, n_items)
        results[rank] = results[rank] and bool(torch.equal(gathered[:, 0], torch.arange(n_items, dtype=torch.float32)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 7, 1])
def test_shard_and_gather_world2(n_items):
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_items, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}


def test_shard_range_properties():
    from highres_net_b200.distributed import shard_range
    for n in (0, 1, 5, 32, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
