"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

Port of the reference's hot path onto the SAME ATen operators the reference
calls (F.interpolate, matmul, F.grid_sample 2-D/3-D, cat, conv1d), so that
(a) it is the CPU baseline timed beside the GPU numbers (bench.py
`cpu_baseline`, `--impl reference`; kind = "port" because the Python reference
cannot travel to the GPU box), and (b) it is the canonical fp32 oracle for the
parity tests.  Pinned against the imported reference by oracle/make_golden.py
(tests/golden/*.npz) and re-checked by tests/test_oracle.py.

Reference lines followed: network/modules.py:24-54 (PerceptualPooling.forward),
:205-214 + :255-282 (VoxelDecoder2), network/models.py:91-97 (LIST.forward query
section), network/executors.py:191-231 (grid loop), utils.py:84-95 (grid order).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .list_oracle import DISPLACEMENT, MAP_SIZE, create_grid_points_from_bounds


def perceptual_pooling(maps: Sequence[torch.Tensor], q: torch.Tensor, trans_mat: torch.Tensor,
                       map_size: int = MAP_SIZE) -> torch.Tensor:
    """modules.py:24-54 -> (B, sum C_i, 1, N)."""
    up = [F.interpolate(m, size=map_size, mode="bilinear", align_corners=True) for m in maps]
    ones = torch.ones(q.shape[0], q.shape[1], 1, device=q.device, dtype=q.dtype)
    h = torch.matmul(torch.cat((q, ones), dim=-1), trans_mat)
    xy = torch.clamp(torch.div(h[:, :, :2], h[:, :, 2:] + 1e-8), 0.0, 136.0)
    half = (map_size - 1) / 2.0
    grid = ((xy - half) / half).unsqueeze(1)
    return torch.cat([F.grid_sample(u, grid, align_corners=True) for u in up], dim=1)


def _displacement_table(device, dtype) -> torch.Tensor:
    rows = [[0.0, 0.0, 0.0]]
    for axis in range(3):
        for sign in (-1, 1):
            r = [0.0, 0.0, 0.0]
            r[axis] = sign * DISPLACEMENT
            rows.append(r)
    return torch.tensor(rows, device=device, dtype=dtype)


def voxel_decoder2(q: torch.Tensor, vols: Sequence[torch.Tensor], percep: torch.Tensor, w: dict) -> torch.Tensor:
    """modules.py:255-282 -> (B, N).  `w`: sdf_decoder state_dict ('fc.fc_0.weight' ...)."""
    table = _displacement_table(q.device, q.dtype)
    q_feat = q.transpose(1, -1)
    qd = torch.cat([q.unsqueeze(1).unsqueeze(1) + d for d in table], dim=2)      # (B,1,7,N,3)
    feats = torch.cat([F.grid_sample(v, qd, padding_mode="border", align_corners=True) for v in vols], dim=1)
    s = feats.shape
    x = torch.cat((feats.reshape(s[0], s[1] * s[3], s[4]), percep, q_feat), dim=1)
    for name in ("fc_0", "fc_1", "fc_2"):
        x = F.relu(F.conv1d(x, w[f"fc.{name}.weight"], w[f"fc.{name}.bias"]))
    return F.conv1d(x, w["fc.fc_out.weight"], w["fc.fc_out.bias"]).squeeze(1)


def list_query(maps, vols, trans_mat, points, weights) -> torch.Tensor:
    """models.py:91-97: raw points (B,N,3) -> scaled SDF (B,N)."""
    B, N, _ = points.shape
    q = points[:, :, [2, 1, 0]] * 2
    percep = perceptual_pooling(maps, q, trans_mat).reshape(B, -1, N)
    return voxel_decoder2(q, vols, percep, weights)


def dense_grid_sdf(maps, vols, trans_mat, weights, res: int, sdf_scale: float,
                   chunk: int = 65536, max_chunks: int | None = None) -> np.ndarray:
    """executors.py:191-231.  `max_chunks` bounds the work for the timed CPU baseline."""
    grid = torch.tensor(create_grid_points_from_bounds(-0.5, 0.5, res)).unsqueeze(0).float()
    out = []
    with torch.no_grad():
        for i, pts in enumerate(torch.split(grid, chunk, 1)):
            if max_chunks is not None and i >= max_chunks:
                break
            out.append(list_query(maps, vols, trans_mat, pts, weights))
    vals = torch.cat(out, dim=1)
    if max_chunks is None:
        vals = vals.view(res, res, res)
    return vals.numpy() / sdf_scale
