"""Mirror of the reference's `network.models` (models.py:14-112): `CoarseNet` and `LIST`.

`LIST` keeps the reference's constructor, attribute names (im_encoder, im_encoder2, point_decoder,
point_mlp_coarse, spatial_transformer, create_occ, vox_encoder, percep_pooling, sdf_decoder),
state_dict keys and `forward(img, query, trans_mat=None) -> (vox_feat[0], sdf)`.
New, hot-path-oriented API on top:
    ctx = model.encode(img[, trans_mat])      once per image  (stock PyTorch encoders + prep kernels)
    sdf = model.query(ctx, points)            per query batch (gather + MLP kernels)
    grid = model.grid(ctx, res, ...)          dense grid / a rank's shard of it
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import hotpath
from . import modules as M


class CoarseNet(nn.Module):
    """RGB image -> coarse point cloud (reference models.py:14-35); stage-1 model, no per-query path."""

    def __init__(self, config):
        super().__init__()
        self.image_encoder = M.ResEncoder()
        self.point_decoder = M.TreeGraphDecoder(config.train_batch_size, config.point_feat, config.point_degree, 10)

    def forward(self, rgba):
        code, _ = self.image_encoder(rgba)
        return self.point_decoder([code.unsqueeze(1)])


class LIST(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.vox_res = config.vox_res
        self.bb_min = getattr(config, "bb_min", -0.5)
        self.bb_max = getattr(config, "bb_max", 0.5)
        #: "fp32" (parity mode, FFMA MLP) or "bf16" (tcgen05 tensor-core MLP); inference only
        self.compute_dtype = getattr(config, "compute_dtype", "fp32")
        enc_feat_size = sum(config.im_enc_layers[3:]) * 7 + 1024 + 3          # models.py:43
        self.vox_encoder = M.VoxelEncoder2(config.im_enc_layers)
        self.sdf_decoder = M.VoxelDecoder2(enc_feat_size, 256)
        self.percep_pooling = M.PerceptualPooling()
        self.im_encoder = M.ResEncoder()
        self.im_encoder2 = M.ResEncoder()
        self.point_decoder = M.TreeGraphDecoder(config.train_batch_size, config.point_feat, config.point_degree, 10)
        self.point_mlp_coarse = M.PointMLP()
        self.spatial_transformer = nn.Sequential(
            nn.Linear(128 + 512, 128), nn.LeakyReLU(0.2), nn.BatchNorm1d(128),
            nn.Linear(128, 128), nn.LeakyReLU(0.2), nn.BatchNorm1d(128),
            nn.Linear(128, 12))

    # ------------------------------------------------------------------ per-image stage (stock PyTorch)
    def per_image(self, img: torch.Tensor, trans_mat: Optional[torch.Tensor] = None, unsqueeze_dim: int = 1):
        """models.py:76-89: returns (feature maps, voxel volumes, trans_mat).
        `unsqueeze_dim=0` reproduces the executor's B=1-only call (executors.py:202)."""
        B = img.shape[0]
        feat_g, _ = self.im_encoder(img)
        feat_g2, feat_l2 = self.im_encoder2(img)
        pc = self.point_decoder([feat_g.unsqueeze(unsqueeze_dim)])
        feat_coarse = self.point_mlp_coarse(pc)
        feat_coarse = torch.max(feat_coarse, -1)[0].reshape(B, -1)
        feat = torch.cat([feat_coarse, feat_g2.reshape(B, -1)], dim=1)
        if trans_mat is None:
            trans_mat = self.spatial_transformer(feat).reshape(-1, 4, 3)
        vox_feat = self.vox_encoder(self.create_occ(pc))
        return feat_l2, vox_feat, trans_mat

    def create_occ(self, pc: torch.Tensor) -> torch.Tensor:
        """models.py:102-112 voxelises the (detached) cloud by nearest grid vertex through a host
        cKDTree; nearest vertex of a regular grid == rounding, so this stays on the device."""
        R = self.vox_res
        p = pc.detach()
        idx = torch.round((p - self.bb_min) / (self.bb_max - self.bb_min) * (R - 1)).clamp_(0, R - 1).long()
        flat = (idx[..., 0] * R + idx[..., 1]) * R + idx[..., 2]
        occ = torch.zeros(p.shape[0], R ** 3, dtype=torch.float, device=p.device)
        occ.scatter_(1, flat, 1.0)
        return occ.view(p.shape[0], R, R, R)

    # ------------------------------------------------------------------ hot path
    @torch.no_grad()
    def encode(self, img: torch.Tensor, trans_mat: Optional[torch.Tensor] = None, dtype: Optional[str] = None,
               unsqueeze_dim: int = 1):
        maps, vols, T = self.per_image(img, trans_mat, unsqueeze_dim)
        ctx = hotpath.prepare_context(maps, vols, T, dtype or self.compute_dtype)
        ctx.occ_pred = vols[0]
        return ctx

    def _weights(self, ctx):
        return self.sdf_decoder.kernel_weights(ctx.layout, "bf16" if ctx.dtype == hotpath._C.BF16 else "fp32")

    @torch.no_grad()
    def query(self, ctx, points: torch.Tensor, chunk_rows: int = hotpath.DEFAULT_CHUNK) -> torch.Tensor:
        """points: raw (B,N,3) in [-0.5,0.5]; the [2,1,0] swap and *2 (models.py:91-92) happen in-kernel."""
        return hotpath.query_sdf(ctx, self._weights(ctx), points, raw=True, chunk_rows=chunk_rows)

    @torch.no_grad()
    def grid(self, ctx, res: int, begin: int = 0, count: Optional[int] = None, sdf_scale: float = 1.0,
             chunk_rows: int = hotpath.DEFAULT_CHUNK) -> torch.Tensor:
        return hotpath.grid_sdf(ctx, self._weights(ctx), res, begin, count, sdf_scale, chunk_rows,
                                self.bb_min, self.bb_max)

    def forward(self, img, query, trans_mat=None):
        """models.py:73-100.  Differentiable: the hot path is one autograd node (fp32 kernels)."""
        maps, vols, T = self.per_image(img, trans_mat)
        if not torch.is_grad_enabled():
            ctx = hotpath.prepare_context(maps, vols, T, self.compute_dtype)
            return vols[0], self.query(ctx, query)
        # kernel layouts through differentiable prep functions (one pass each, no torch.cat / permute copies), so
        # autograd reaches the encoders
        maps_cl = hotpath.prep_maps_autograd(maps, self.percep_pooling.map_size)
        vols_cl = [hotpath.prep_volume_autograd(v) for v in vols]
        sdf = hotpath.query_sdf_autograd(query, T, maps_cl, vols_cl, self.sdf_decoder.param_dict(), raw=True)
        return vols[0], sdf
