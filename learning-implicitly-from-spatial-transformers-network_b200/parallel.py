"""Multi-GPU sharding of the dense grid: one process per GPU, contiguous ranges
of the flattened res^3 grid per rank (x-slabs), ONE all_gather of fp32 SDF values at the end.
The reference's only parallelism is nn.DataParallel (train.py:126, test.py:62), a no-op at test
time; nothing in rows a-2..a-6 mixes query points, so the partition is exact."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


COARSE_MAX_RES = 32      # levels up to this resolution are read by the projection (csrc/api.cu kLinesMaxRes) and cannot be late


def shard_range(total: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """[begin, count) of `rank`: equal contiguous ranges, boundaries rounded to `align`
    (the last rank takes the remainder)."""
    per = -(-total // world)
    per = -(-per // align) * align
    begin = min(total, rank * per)
    return begin, max(0, min(total, begin + per) - begin)


def gather_shards(local: torch.Tensor, total: int, world: int, align: int = 1, group=None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """local: (B, count_r) fp32 shard of every image -> (B, total) on every rank, ONE collective (all_gather of equal
    pieces of `per` = the aligned shard size).  When `local` already is a (B, per) buffer -- the caller let the kernels
    write straight into it (shard_buffer) -- nothing is staged; `out` (world * B, per) is reused if given."""
    if world == 1:
        return local
    B = local.shape[0]
    per = shard_range(total, 0, world, align)[1]
    if local.shape[1] == per and local.is_contiguous():
        buf = local
    else:
        buf = torch.zeros(B, per, device=local.device, dtype=local.dtype)
        buf[:, :local.shape[1]] = local
    if out is None:
        out = torch.empty(world * B, per, device=local.device, dtype=local.dtype)      # rank-major concat on dim 0
    dist.all_gather_into_tensor(out, buf, group=group)
    if B == 1 and world * per == total:
        return out.view(1, total)                                                       # already the grid, no copy
    return out.view(world, B, per).permute(1, 0, 2).reshape(B, world * per)[:, :total].contiguous()


def shard_buffer(total: int, rank: int, world: int, B: int, device, align: int = 1):
    """(buffer (B, per), view (B, count_r)): let the evaluation write its shard into `view`; `buffer` then goes into
    gather_shards without a staging copy (the tail of the last rank's buffer stays zero)."""
    per = shard_range(total, 0, world, align)[1]
    count = shard_range(total, rank, world, align)[1]
    buf = torch.zeros(B, per, device=device, dtype=torch.float32)
    return buf, buf[:, :count]


def sharded_grid(evaluate: Callable[[int, int], torch.Tensor], total: int, align: int = 1, group=None) -> torch.Tensor:
    """evaluate(begin, count) -> (B, count); returns the gathered (B, total) grid on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return evaluate(0, total)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    begin, count = shard_range(total, rank, world, align)
    return gather_shards(evaluate(begin, count), total, world, align, group)


def slice_bounds(numel: int, per: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of `rank`'s slice of a flattened tensor cut into pieces of `per` elements."""
    lo = min(numel, rank * per)
    return lo, min(numel, lo + per)


def rebuild_from_gathered(gathered: torch.Tensor, per, fulls) -> None:
    """gathered: (world, sum(per)) rank-major slices of every tensor (each padded to per[i]); writes the
    reassembled tensors into `fulls` in place."""
    off = 0
    for d, p in zip(fulls, per):
        d.view(-1).copy_(gathered[:, off:off + p].reshape(-1)[:d.numel()])
        off += p


def upload_stages(vol_res, n_maps: int, n_tensors: int):
    """Index lists (into [*maps, *vols, T]) of the two upload stages of the staged end-to-end path: everything the
    projection and the first chunk's line tables read (maps, levels with R <= 32, T), then the fine levels.  Returns
    (stages, late_levels); a stage is dropped if empty."""
    late_levels = [l for l, r in enumerate(vol_res) if r > COARSE_MAX_RES]
    late_idx = {n_maps + l for l in late_levels}
    stages = [[i for i in range(n_tensors) if i not in late_idx], sorted(late_idx)]
    return [st for st in stages if st], late_levels


def gather_stage(hosts, idx, per, mine: torch.Tensor, gathered: torch.Tensor, fulls, rank: int, group=None) -> None:
    """One upload stage: this rank's 1/world slice of every tensor in `idx` -> `mine` (device), ONE all_gather,
    reassembly into fulls[i].  Device agnostic (the gloo tests run it on the CPU)."""
    off = 0
    for i, p in zip(idx, per):
        flat = hosts[i].view(-1)
        lo, hi = slice_bounds(flat.numel(), p, rank)
        if hi > lo:
            mine[off:off + hi - lo].copy_(flat[lo:hi], non_blocking=True)
        off += p
    dist.all_gather_into_tensor(gathered.view(-1), mine, group=group)
    rebuild_from_gathered(gathered, per, [fulls[i] for i in idx])


class ShardedHostRunner:
    """End-to-end dense-grid evaluation from HOST buffers on `world` ranks (bench.py's e2e leg):
    pinned reference-layout per-image tensors -> device -> prep kernels -> this rank's shard of the grid -> pinned
    host shard.  With one rank this is the C-ABI host-buffer call (list_sdf_grid_host), which overlaps the upload of
    the big volumes with the projection and the first chunk's addend gather, and every chunk's download with the next
    chunk's kernels.  With several ranks, uploading the same 206 MB on every rank would make the step PCIe-bound, so
    each rank uploads 1/world of every tensor and ONE all_gather over NVLink rebuilds the full set on every GPU (the
    per-image tensors are replicated, the grid is what is sharded)."""

    def __init__(self, maps_host, vols_host, trans_host, weights, res: int, dtype="bf16", chunk_rows: int = 1048576,
                 group=None):
        from . import hotpath
        self.hotpath = hotpath
        self.group = group
        ddp = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if ddp else 0
        self.world = dist.get_world_size(group) if ddp else 1
        self.hosts = [*maps_host, *vols_host, trans_host]
        for t in self.hosts:
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or not t.is_pinned():
                raise ValueError("host inputs must be contiguous, pinned fp32 CPU tensors")
        self.n_maps = len(maps_host)
        self.weights, self.res, self.dtype = weights, res, dtype
        dev = weights.w0.device
        self.dev = dev
        total = res ** 3
        self.begin, self.count = shard_range(total, self.rank, self.world, align=res * res)
        self.chunk = max(1, min(chunk_rows, self.count))
        self.single = None
        if self.world == 1:
            self.single = hotpath.HostGridRunner(maps_host, vols_host, trans_host, weights, res, self.begin, self.count,
                                                 dtype, self.chunk)
            self.h2d_bytes, self.d2h_bytes = self.single.h2d_bytes, self.single.d2h_bytes
            return
        self.full = [torch.empty(t.shape, device=dev, dtype=torch.float32) for t in self.hosts]
        self.out_dev = torch.empty(trans_host.shape[0], self.count, device=dev, dtype=torch.float32)
        self.out_host = torch.empty(trans_host.shape[0], self.count, dtype=torch.float32).pin_memory()
        self.per = [-(-t.numel() // self.world) for t in self.hosts]           # elements of every rank's slice
        # Two stages (same idea as list_sdf_grid_host): the coarse tensors -- maps, levels with R <= 32, T,
        # all the projection and the first addend gather read -- are uploaded and all-gathered first; the fine levels
        # follow on a side stream while those kernels run, and list_sdf_grid_late prepares them when they are there.
        stage_idx, self.late_levels = upload_stages([v.shape[2] for v in vols_host], self.n_maps, len(self.hosts))
        self.stages = []
        for idx in stage_idx:
            per = [self.per[i] for i in idx]
            self.stages.append({"idx": idx, "per": per,
                                "mine": torch.zeros(sum(per), device=dev, dtype=torch.float32),
                                "gathered": torch.empty(self.world, sum(per), device=dev, dtype=torch.float32),
                                "event": torch.cuda.Event()})
        self.side = torch.cuda.Stream(device=dev)
        self.h2d_bytes = sum((slice_bounds(t.numel(), p, self.rank)[1] - slice_bounds(t.numel(), p, self.rank)[0]) * 4
                             for t, p in zip(self.hosts, self.per))
        self.d2h_bytes = self.out_host.numel() * 4
        self.workspace = None

    def run(self, sdf_scale: float = 1.0) -> torch.Tensor:
        hp = self.hotpath
        if self.single is not None:
            return self.single.run(sdf_scale)
        main = torch.cuda.current_stream(self.dev)
        self.side.wait_stream(main)                      # the previous call's readers of self.full are done
        with torch.cuda.stream(self.side):
            for st in self.stages:
                gather_stage(self.hosts, st["idx"], st["per"], st["mine"], st["gathered"], self.full, self.rank, self.group)
                st["event"].record(self.side)
        main.wait_event(self.stages[0]["event"])
        maps, vols, T = self.full[:self.n_maps], self.full[self.n_maps:-1], self.full[-1]
        staged = len(self.stages) > 1
        ctx = hp.prepare_context(maps, vols, T, self.dtype, skip_levels=self.late_levels if staged else ())
        if self.workspace is None:
            self.workspace = hp._workspace(ctx.struct(), self.weights.struct(), self.chunk, self.dev, self.res)
        late = [vols[l] if (staged and l in self.late_levels) else None for l in range(len(vols))]
        hp.grid_sdf_late(ctx, self.weights, self.res, self.begin, self.count, sdf_scale, self.chunk, self.out_dev,
                         self.workspace, self.stages[-1]["event"] if staged else None, late, out_host=self.out_host)
        return self.out_host
