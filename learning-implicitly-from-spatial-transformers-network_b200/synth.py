"""Seeded synthetic inputs for the LIST per-query hot path.

Everything the hot path consumes is produced here from a seed on the CPU with
``torch.Generator`` so that the GPU box (which has no /root/reference) can
regenerate bit-identical inputs for the committed golden fixtures
(tests/golden/, written by oracle/make_golden.py).

Shapes follow what the reference's per-image stage emits for one 224x224
image (probed): five ResNet-18 feature maps
``64x224^2, 64x112^2, 128x56^2, 256x28^2, 512x14^2`` (reference
``network/modules.py:1050-1074``) and six voxel-encoder volumes
``1x128^3, 16x128^3, 32x64^3, 64x32^3, 128x16^3, 128x8^3`` (reference
``network/modules.py:425-442``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch

# (channels, spatial size)
MAP_SHAPES_FULL = [(64, 224), (64, 112), (128, 56), (256, 28), (512, 14)]
VOL_SHAPES_FULL = [(1, 128), (16, 128), (32, 64), (64, 32), (128, 16), (128, 8)]
# same channel counts (the kernels are specialised on them), small extents
MAP_SHAPES_SMALL = [(64, 24), (64, 12), (128, 9), (256, 5), (512, 3)]
VOL_SHAPES_SMALL = [(1, 16), (16, 16), (32, 8), (64, 6), (128, 4), (128, 2)]

MAP_SIZE = 137          # reference PerceptualPooling(map_size=137), modules.py:16
H_DIM = 256             # reference models.py:47
SEED = 333              # reference train.py:18


@dataclass
class HotPathInputs:
    maps: List[torch.Tensor]            # 5 x (B, C_i, H_i, W_i) fp32, NCHW
    vols: List[torch.Tensor]            # 6 x (B, C, R, R, R)   fp32, NCDHW
    trans_mat: torch.Tensor             # (B, 4, 3)
    points: torch.Tensor                # (B, N, 3) raw query points in [-0.5, 0.5] (x, y, z)
    weights: dict = field(default_factory=dict)   # reference state_dict keys of sdf_decoder.fc.*

    @property
    def feature_size(self) -> int:
        return sum(v.shape[1] for v in self.vols) * 7 + sum(m.shape[1] for m in self.maps) + 3

    def to(self, device) -> "HotPathInputs":
        return HotPathInputs(
            [m.to(device) for m in self.maps],
            [v.to(device) for v in self.vols],
            self.trans_mat.to(device),
            self.points.to(device),
            {k: v.to(device) for k, v in self.weights.items()},
        )


def camera_like_transmat(B: int, gen: torch.Generator, noise: float = 0.05) -> torch.Tensor:
    """A 4x3 matrix that keeps <~5 % of queries on the clamp."""
    base = torch.tensor(
        [[60.0, 0.0, 0.20],
         [0.0, 60.0, 0.10],
         [5.0, -5.0, 0.30],
         [136.0, 136.0, 2.0]]
    )
    T = base.unsqueeze(0).repeat(B, 1, 1)
    return T * (1.0 + noise * (torch.rand(T.shape, generator=gen) - 0.5))


def random_transmat(B: int, gen: torch.Generator) -> torch.Tensor:
    """What a random-init spatial_transformer emits: O(0.1) entries of both signs,
    so ~half the queries clamp and the divide's singular plane cuts the grid."""
    return (torch.rand(B, 4, 3, generator=gen) - 0.5) * 0.6


def mlp_weights(feature_size: int, gen: torch.Generator, h_dim: int = H_DIM) -> dict:
    """Conv1d default init U(+-1/sqrt(fan_in)) under the reference's state_dict
    names (reference modules.py:196-200)."""
    dims = [("fc_0", feature_size, 2 * h_dim), ("fc_1", 2 * h_dim, h_dim),
            ("fc_2", h_dim, h_dim), ("fc_out", h_dim, 1)]
    out = {}
    for name, fin, fout in dims:
        bound = 1.0 / math.sqrt(fin)
        out[f"fc.{name}.weight"] = (torch.rand(fout, fin, 1, generator=gen) * 2 - 1) * bound
        out[f"fc.{name}.bias"] = (torch.rand(fout, generator=gen) * 2 - 1) * bound
    return out


def training_points(B: int, N: int, gen: torch.Generator,
                    distribution=(0.45, 0.44, 0.10), sigmas=(0.003, 0.01, 0.07),
                    radius: float = 0.35):
    """cfg-2 query recipe (reference datasets/Datasets.py:153-154,
    219-230): surface samples of a sphere + sigma-perturbation, analytic SDF."""
    counts = [int(round(d * N)) for d in distribution]
    counts[0] += N - sum(counts)
    chunks = []
    for c, s in zip(counts, sigmas):
        d = torch.randn(B, c, 3, generator=gen)
        d = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-12) * radius
        chunks.append(d + s * torch.randn(B, c, 3, generator=gen))
    pts = torch.cat(chunks, dim=1).clamp(-0.5, 0.5)
    sdf = pts.norm(dim=-1) - radius
    return pts, sdf


def make_inputs(seed: int = SEED, B: int = 1, N: int = 2048, size: str = "full",
                trans: str = "camera", points: str = "uniform",
                map_shapes: Optional[list] = None, vol_shapes: Optional[list] = None) -> HotPathInputs:
    gen = torch.Generator().manual_seed(seed)
    ms = map_shapes or (MAP_SHAPES_FULL if size == "full" else MAP_SHAPES_SMALL)
    vs = vol_shapes or (VOL_SHAPES_FULL if size == "full" else VOL_SHAPES_SMALL)
    maps = [torch.randn(B, c, s, s, generator=gen).relu_() for c, s in ms]
    vols = []
    for i, (c, r) in enumerate(vs):
        if i == 0:   # sigmoid occupancy in (0,1) (reference modules.py:432-434)
            vols.append(torch.rand(B, c, r, r, r, generator=gen))
        else:
            vols.append(torch.randn(B, c, r, r, r, generator=gen))
    T = camera_like_transmat(B, gen) if trans == "camera" else random_transmat(B, gen)
    if points == "uniform":
        pts = torch.rand(B, N, 3, generator=gen) - 0.5
    elif points == "training":
        pts, _ = training_points(B, N, gen)
    else:
        raise ValueError(points)
    inp = HotPathInputs(maps, vols, T, pts)
    inp.weights = mlp_weights(inp.feature_size, gen)
    return inp
