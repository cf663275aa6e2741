"""CPU tests of the N>1 path: shard arithmetic and the single gather, with
world_size-2 gloo processes."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from list_b200 import parallel


def test_shard_range_covers_exactly():
    for total in (0, 1, 7, 64 ** 3, 256 ** 3, 1000):
        for world in (1, 2, 3, 4, 8):
            for align in (1, 64, 65536):
                spans = [parallel.shard_range(total, r, world, align) for r in range(world)]
                pos = 0
                for b, c in spans:
                    assert b == min(pos, total) and c >= 0
                    pos = b + c
                assert pos == total
                assert all(b % align == 0 for b, c in spans if c > 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, align, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def evaluate(begin, count):          # stands in for the per-rank kernel launch
            idx = torch.arange(begin, begin + count, dtype=torch.float32)
            return torch.stack([idx * 2 + 1, -idx])
        out = parallel.sharded_grid(evaluate, total, align)
        ref = torch.arange(total, dtype=torch.float32)
        ok = torch.equal(out, torch.stack([ref * 2 + 1, -ref]))
        q.put((rank, bool(ok), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total,align", [(4096, 256), (1000, 1), (27, 9)])
def test_sharded_grid_gloo_world2(total, align):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, align, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (2, total) for _, _, shape in res)


def test_sliced_upload_reassembles_every_tensor():
    """ShardedHostRunner's H2D scheme: every rank uploads 1/world of each tensor, one all_gather, then reassembly."""
    import torch
    from list_b200 import parallel
    g = torch.Generator().manual_seed(0)
    tensors = [torch.rand(3, 5, 7, generator=g), torch.rand(11, generator=g), torch.rand(1, 4, 3, generator=g), torch.rand(2, 2, generator=g)]
    for world in (1, 2, 3, 8):
        per = [-(-t.numel() // world) for t in tensors]
        gathered = torch.zeros(world, sum(per))
        for rank in range(world):
            off = 0
            for t, p in zip(tensors, per):
                lo, hi = parallel.slice_bounds(t.numel(), p, rank)
                gathered[rank, off:off + hi - lo] = t.view(-1)[lo:hi]
                off += p
        fulls = [torch.empty_like(t) for t in tensors]
        parallel.rebuild_from_gathered(gathered, per, fulls)
        for a, b in zip(fulls, tensors):
            assert torch.equal(a, b)


def _stage_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)                      # same tensors on every rank
        maps = [torch.rand(1, 4, 6, 6, generator=g), torch.rand(1, 8, 3, 3, generator=g)]
        vols = [torch.rand(1, 1, 40, 40, 40, generator=g), torch.rand(1, 3, 36, 36, 36, generator=g),
                torch.rand(1, 8, 16, 16, 16, generator=g), torch.rand(1, 8, 5, 5, 5, generator=g)]
        T = torch.rand(1, 4, 3, generator=g)
        hosts = [*maps, *vols, T]
        stages, late = parallel.upload_stages([v.shape[2] for v in vols], len(maps), len(hosts))
        fulls = [torch.full_like(t, float("nan")) for t in hosts]
        per_all = [-(-t.numel() // world) for t in hosts]
        seen = []
        for idx in stages:
            per = [per_all[i] for i in idx]
            mine = torch.zeros(sum(per))
            gathered = torch.empty(world, sum(per))
            parallel.gather_stage(hosts, idx, per, mine, gathered, fulls, rank)
            seen.append([bool(torch.equal(fulls[i], hosts[i])) for i in range(len(hosts))])
        q.put((rank, stages, late, seen))
    finally:
        dist.destroy_process_group()


def test_staged_upload_gloo_world2():
    """The two-stage upload of the multi-rank end-to-end path: after stage 1 exactly the coarse tensors (maps, levels
    with R <= 32, T) are complete on every rank, after stage 2 all of them."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stage_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, stages, late, seen in res:
        assert late == [0, 1] and stages == [[0, 1, 4, 5, 6], [2, 3]], (stages, late)
        assert seen[0] == [True, True, False, False, True, True, True], seen
        assert seen[1] == [True] * 7, seen


def test_upload_stages_without_fine_levels_is_one_stage():
    stages, late = parallel.upload_stages([16, 8, 4], 2, 6)
    assert late == [] and stages == [[0, 1, 2, 3, 4, 5]]
