"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU fp32 restatement of LIST's per-query SDF hot path as explicit index
arithmetic + gathers + matmuls (no grid_sample / interpolate / Conv1d), used to
check the CUDA path stage by stage.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this file.

Parity status: PINNED.  The reference has no tests or golden vectors of its own
(SURVEY.md §4, §8c), so this restatement (and oracle/ref_port.py) is pinned
against outputs of the reference itself, imported unmodified from
/root/reference by oracle/make_golden.py, and stored in tests/golden/*.npz;
tests/test_oracle.py re-checks that on every run.

Each function cites the reference lines (relative to /root/reference) and the
torch 2.11 ATen arithmetic it follows (the arithmetic lives in
ATen: ATen/native/UpSample.h, ATen/native/GridSampler.h).
Everything is written with differentiable torch ops so torch.autograd over this
file is the oracle for the backward path (row a-9) as well.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np
import torch

DISPLACEMENT = 0.0722      # reference network/modules.py:205
MAP_SIZE = 137             # reference network/modules.py:16


# --------------------------------------------------------------------------- a-1
def _upsample_axis(in_size: int, out_size: int):
    """ATen area_pixel_compute_scale + guard_index_and_lambda, align_corners=True
    (UpSample.h:271-299, 442-448): src = dst*(in-1)/(out-1) in fp32."""
    scale = np.float32(in_size - 1) / np.float32(out_size - 1) if out_size > 1 else np.float32(0)
    dst = torch.arange(out_size, dtype=torch.float32)
    src = dst * float(scale)
    i0 = src.floor().to(torch.int64).clamp(max=in_size - 1)
    lam = (src - i0.to(torch.float32)).clamp(0.0, 1.0)
    i1 = i0 + (i0 < in_size - 1).to(torch.int64)
    return i0, i1, lam


def upsample_bilinear_align_corners(x: torch.Tensor, size: int = MAP_SIZE) -> torch.Tensor:
    """reference modules.py:26-35: F.interpolate(x, size, 'bilinear', align_corners=True).
    x: (B, C, H, W) -> (B, C, size, size)."""
    _, _, H, W = x.shape
    y0, y1, ly = _upsample_axis(H, size)
    x0, x1, lx = _upsample_axis(W, size)
    ly = ly.view(1, 1, size, 1)
    lx = lx.view(1, 1, 1, size)
    r0 = x[:, :, y0, :]
    r1 = x[:, :, y1, :]
    top = (1.0 - lx) * r0[:, :, :, x0] + lx * r0[:, :, :, x1]
    bot = (1.0 - lx) * r1[:, :, :, x0] + lx * r1[:, :, :, x1]
    return (1.0 - ly) * top + ly * bot


def prepare_maps(maps: Sequence[torch.Tensor], size: int = MAP_SIZE) -> torch.Tensor:
    """a-1 hoisted: 5 maps -> one channels-last (B, size, size, sum C) tensor, channel order
    f1..f5 as in the reference's cat (modules.py:53)."""
    ups = [upsample_bilinear_align_corners(m, size) for m in maps]
    return torch.cat(ups, dim=1).permute(0, 2, 3, 1).contiguous()


# --------------------------------------------------------------------------- a-2
def _fma32(a, b, c):
    """fp32 fused multiply-add emulated in fp64 (a*b exact in fp64)."""
    return (a.double() * b.double() + c.double()).float()


def localise(q: torch.Tensor, trans_mat: torch.Tensor, fused: bool = False):
    """reference modules.py:37-43.  q: (B,N,3) already swapped+scaled, trans_mat (B,4,3).
    h = [q,1]·T accumulated k=0..3 (sgemm order); xy = h[:2]/(h[2]+1e-8); clamp [0,136].
    `fused=True` uses FMA accumulation exactly as the CUDA kernel does (non-differentiable)."""
    T = trans_mat
    if fused:
        acc = q[..., 0:1] * T[:, None, 0, :]
        acc = _fma32(q[..., 1:2], T[:, None, 1, :], acc)
        acc = _fma32(q[..., 2:3], T[:, None, 2, :], acc)
        h = acc + T[:, None, 3, :]
    else:
        h = q[..., 0:1] * T[:, None, 0, :] + q[..., 1:2] * T[:, None, 1, :] \
            + q[..., 2:3] * T[:, None, 2, :] + T[:, None, 3, :]
    xy = h[..., :2] / (h[..., 2:] + 1e-8)
    lim = float(MAP_SIZE - 1)
    # torch.clamp propagates NaN; clamped coordinates get zero gradient (autograd of clamp)
    return torch.clamp(xy, 0.0, lim), h


# --------------------------------------------------------------------------- a-3
def gather2d(maps_cl: torch.Tensor, xy: torch.Tensor) -> torch.Tensor:
    """reference modules.py:45-53 = grid_sample(bilinear, zeros, align_corners=True) on the
    channels-last upsampled maps.  maps_cl: (B,S,S,C); xy: (B,N,2) in pixels, xy[...,0]->W.
    Returns (B,N,C).  ATen GridSampler.h: ix=((g+1)/2)*(S-1); taps outside [0,S-1] add 0."""
    B, S, _, C = maps_cl.shape
    half = (S - 1) / 2.0
    g = (xy - half) / half
    ix = ((g[..., 0] + 1.0) / 2.0) * (S - 1)
    iy = ((g[..., 1] + 1.0) / 2.0) * (S - 1)
    nan = torch.isnan(ix) | torch.isnan(iy)          # non-finite grid -> all taps out of bounds
    ix = torch.where(nan, torch.zeros_like(ix), ix)
    iy = torch.where(nan, torch.zeros_like(iy), iy)
    x0 = ix.floor()
    y0 = iy.floor()
    x1 = x0 + 1
    y1 = y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    bidx = torch.arange(B).view(B, 1).expand_as(ix)

    def tap(xi, yi, w):
        ok = (xi >= 0) & (xi <= S - 1) & (yi >= 0) & (yi <= S - 1) & ~nan
        xc = xi.clamp(0, S - 1).long()
        yc = yi.clamp(0, S - 1).long()
        v = maps_cl[bidx, yc, xc]                      # (B,N,C)
        return v * (w * ok.to(w.dtype)).unsqueeze(-1)

    return tap(x0, y0, w_nw) + tap(x1, y0, w_ne) + tap(x0, y1, w_sw) + tap(x1, y1, w_se)


# --------------------------------------------------------------------------- a-4
def displacements(dtype=torch.float32) -> torch.Tensor:
    """reference modules.py:205-214: [0,0,0] then for axis in x,y,z: -d, +d."""
    rows = [[0.0, 0.0, 0.0]]
    for axis in range(3):
        for sign in (-1.0, 1.0):
            r = [0.0, 0.0, 0.0]
            r[axis] = sign * DISPLACEMENT
            rows.append(r)
    return torch.tensor(rows, dtype=dtype)


# --------------------------------------------------------------------------- a-5
def gather3d(vol_cl: torch.Tensor, pts: torch.Tensor) -> torch.Tensor:
    """grid_sample(trilinear, border, align_corners=True) (reference modules.py:264-265) on a
    channels-last volume.  vol_cl: (B,D,H,W,C); pts: (B,M,3) with pts[...,0]->W, [1]->H, [2]->D.
    ATen: i = clamp(((c+1)/2)*(R-1), 0, R-1); corners floor / floor+1, 8 taps summed in the
    order tnw,tne,tsw,tse,bnw,bne,bsw,bse; a corner index == R is skipped (its weight is 0)."""
    B, D, H, W, C = vol_cl.shape

    def unnorm(c, size):
        return (((c + 1.0) / 2.0) * (size - 1)).clamp(0.0, float(size - 1))

    ix, iy, iz = unnorm(pts[..., 0], W), unnorm(pts[..., 1], H), unnorm(pts[..., 2], D)
    x0, y0, z0 = ix.floor(), iy.floor(), iz.floor()
    x1, y1, z1 = x0 + 1, y0 + 1, z0 + 1
    wx0, wx1 = x1 - ix, ix - x0
    wy0, wy1 = y1 - iy, iy - y0
    wz0, wz1 = z1 - iz, iz - z0
    bidx = torch.arange(B).view(B, 1).expand_as(ix)
    out = 0
    for (zi, wz) in ((z0, wz0), (z1, wz1)):
        for (yi, wy) in ((y0, wy0), (y1, wy1)):
            for (xi, wx) in ((x0, wx0), (x1, wx1)):
                ok = (xi <= W - 1) & (yi <= H - 1) & (zi <= D - 1)
                v = vol_cl[bidx, zi.clamp(max=D - 1).long(), yi.clamp(max=H - 1).long(),
                           xi.clamp(max=W - 1).long()]
                out = out + v * ((wx * wy * wz) * ok.to(wx.dtype)).unsqueeze(-1)
    return out


def voxel_features(vols_cl: Sequence[torch.Tensor], q: torch.Tensor) -> torch.Tensor:
    """reference modules.py:256-273: 7 displaced copies, 6 trilinear gathers, cat over volumes,
    reshape so that feature index = c_global*7 + d.  Returns (B,N,7*sumC)."""
    B, N, _ = q.shape
    disp = displacements(q.dtype).to(q.device)
    qd = (q.unsqueeze(1) + disp.view(1, 7, 1, 3)).reshape(B, 7 * N, 3)       # d-major
    feats = [gather3d(v, qd).reshape(B, 7, N, v.shape[-1]) for v in vols_cl]
    f = torch.cat(feats, dim=-1)                                             # (B,7,N,sumC)
    return f.permute(0, 2, 3, 1).reshape(B, N, -1)                           # index c*7+d


# --------------------------------------------------------------------------- a-6
def implicit_mlp(x: torch.Tensor, w: dict) -> torch.Tensor:
    """reference modules.py:196-201, 276-282.  x: (B,N,K) -> (B,N).  `w` uses the
    reference's state_dict names relative to sdf_decoder ('fc.fc_0.weight' [512,K,1] ...)."""
    def lin(h, name):
        return h @ w[f"fc.{name}.weight"].squeeze(-1).t() + w[f"fc.{name}.bias"]
    h = torch.relu(lin(x, "fc_0"))
    h = torch.relu(lin(h, "fc_1"))
    h = torch.relu(lin(h, "fc_2"))
    return lin(h, "fc_out").squeeze(-1)


# --------------------------------------------------------------------------- a-7
def to_channels_last_vol(v: torch.Tensor) -> torch.Tensor:
    return v.permute(0, 2, 3, 4, 1).contiguous()


def feature_rows(maps: Sequence[torch.Tensor], vols: Sequence[torch.Tensor],
                 trans_mat: torch.Tensor, points: torch.Tensor, fused_localise: bool = False):
    """The (B,N,3610) fc_0 input in the REFERENCE's column order
    [vox(c*7+d) | percep(1024) | q(3)] (modules.py:275), from raw query points."""
    q = points[:, :, [2, 1, 0]] * 2                                          # models.py:91-92
    maps_cl = prepare_maps(maps)
    xy, _ = localise(q, trans_mat, fused=fused_localise)
    percep = gather2d(maps_cl, xy)
    vox = voxel_features([to_channels_last_vol(v) for v in vols], q)
    return torch.cat([vox, percep, q], dim=-1)


def list_query(maps, vols, trans_mat, points, weights, fused_localise: bool = False) -> torch.Tensor:
    """models.py:91-97 end to end: raw query points (B,N,3) in [-0.5,0.5] -> scaled SDF (B,N)."""
    return implicit_mlp(feature_rows(maps, vols, trans_mat, points, fused_localise), weights)


# --------------------------------------------------------------------------- a-8
def create_grid_points_from_bounds(minimum: float, maximum: float, res: int) -> np.ndarray:
    """reference utils.py:84-95: linspace + meshgrid('ij'), x slowest / z fastest, float64."""
    ax = np.linspace(minimum, maximum, res)
    idx = np.arange(res ** 3)
    return np.stack([ax[idx // (res * res)], ax[(idx // res) % res], ax[idx % res]], axis=1)


def dense_grid_sdf(maps, vols, trans_mat, weights, res: int, sdf_scale: float,
                   chunk: int = 65536, bb_min: float = -0.5, bb_max: float = 0.5) -> np.ndarray:
    """reference executors.py:191-231: chunked evaluation of the res^3 grid, /sdf_scale."""
    grid = torch.tensor(create_grid_points_from_bounds(bb_min, bb_max, res)).unsqueeze(0).float()
    out = []
    with torch.no_grad():
        for pts in torch.split(grid, chunk, 1):
            out.append(list_query(maps, vols, trans_mat, pts, weights))
    vals = torch.cat(out, dim=1).view(res, res, res).numpy()
    return vals / sdf_scale


# --------------------------------------------------------------------------- a-9
def sdf_loss(pred: torch.Tensor, target: torch.Tensor, sdf_scale: float) -> torch.Tensor:
    """reference losses.py:21-22: mean_b sum_n (gt*scale - pred)^2."""
    return ((target * sdf_scale - pred) ** 2).sum(-1).mean()


def occ_loss(occ: torch.Tensor, occ_gt: torch.Tensor, w: float = 0.9) -> torch.Tensor:
    """reference executors.py:138-141."""
    return 1000 * (-w * torch.mean(occ_gt * torch.log(occ + 1e-8))
                   - (1 - w) * torch.mean((1 - occ_gt) * torch.log(1 - occ + 1e-8)))


def mc_case_index(grid: np.ndarray) -> np.ndarray:
    """Marching-cubes case index (0..255) of every cell of -grid at threshold 0
    (reference utils.py:172-173 calls mcubes.marching_cubes(-grid, 0)); two fields with equal
    case indices give topologically identical meshes."""
    inside = (-grid) < 0
    idx = np.zeros(tuple(s - 1 for s in grid.shape), dtype=np.uint8)
    bit = 0
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                sl = inside[dx:grid.shape[0] - 1 + dx, dy:grid.shape[1] - 1 + dy, dz:grid.shape[2] - 1 + dz]
                idx |= (sl.astype(np.uint8) << bit)
                bit += 1
    return idx
