"""Mirror of the reference's `network.executors.LIST` (executors.py:102-268): same constructor,
`.train(batch, calc_loss=True) -> (pred, loss_dict)`, `.test(batch, eval_pred=False) ->
([mesh, occ, occ_pred], scores)`, `.eval`, `.save`, `.create_grid`; selected through the
reference's dotted-path plugin lookup (`--model list_b200.network.models.LIST`,
utils.py:20-26, test.py:60,95).

What changed underneath (rows a-8 / §8e):
  * the res^3 grid is generated on the device (no per-chunk H2D), every chunk runs gather + MLP
    kernels back to back on one stream, and there is a single D2H of the finished grid instead of
    a sync per chunk (executors.py:215-224);
  * under torch.distributed the grid is sharded by contiguous point ranges across ranks and
    gathered with one collective;
  * `grid_res` (default = vox_res) decouples the query grid from the voxel resolution, which the
    reference ties together (executors.py:192-193,229).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import hotpath, parallel
from .losses import SDFLoss


class LIST:
    def __init__(self, config, model):
        self.model = model
        self.use_cuda = getattr(config, "cuda", True)
        self.device = getattr(config, "device", "cuda")
        self.test_pointnum = getattr(config, "test_pointnum", 65536)
        self.sdf_scale = getattr(config, "sdf_scale", 1.0)
        self.max_dist = getattr(config, "sdf_max_dist", 1.0)
        self.mcube_znum = getattr(config, "mcube_znum", 128)
        self.bb_min = getattr(config, "bb_min", -0.5)
        self.bb_max = getattr(config, "bb_max", 0.5)
        self.vox_res = config.vox_res
        self.grid_res = getattr(config, "grid_res", None) or config.vox_res
        self.loss_sdf = SDFLoss(self.sdf_scale)

    @property
    def net(self):
        return self.model.module if hasattr(self.model, "module") else self.model

    def create_grid(self):
        """utils.create_grid_points_from_bounds on the device, (res^3, 3) fp32."""
        return hotpath.grid_points(self.grid_res, 0, None, self.bb_min, self.bb_max, self.device)

    def calc_loss(self, pred, gt):
        occ, sdf_pred = pred
        occ_gt, sdf_gt = gt
        w = 0.9                                                                      # executors.py:138-141
        occ_loss = 1000 * (-w * torch.mean(occ_gt * torch.log(occ + 1e-8))
                           - (1 - w) * torch.mean((1 - occ_gt) * torch.log(1 - occ + 1e-8)))
        loss = {"occ_loss": occ_loss}
        loss.update(self.loss_sdf(sdf_pred, sdf_gt))
        return loss

    def train(self, batch, calc_loss=True):
        img, points, sdf_gt, occ_gt = batch["rgb_image"], batch["points"], batch["values"], batch["occ"]
        transmat = batch.get("transmat")
        dev = self.device
        img, points, sdf_gt, occ_gt = img.to(dev), points.to(dev), sdf_gt.to(dev), occ_gt.to(dev)
        if transmat is not None:
            transmat = transmat.to(dev)
        pred = self.model(img, points, transmat)
        loss = self.calc_loss(pred, [occ_gt, sdf_gt]) if calc_loss else []
        return pred, loss

    @torch.no_grad()
    def predict_grid(self, batch):
        """SDF grid (res,res,res) float32 numpy, already divided by sdf_scale (executors.py:226-231)."""
        img = batch["rgb_image"].to(self.device)
        transmat = batch.get("transmat")
        if transmat is not None:
            transmat = transmat.to(self.device)
        net = self.net
        ctx = net.encode(img, transmat, unsqueeze_dim=0 if img.shape[0] == 1 else 1)
        res = self.grid_res
        total = res ** 3
        grid = parallel.sharded_grid(
            # results do not depend on the chunking (tests), so the kernels get large launches instead of the
            # reference's 65 536-point chunks: the last partial wave of gather CTAs is amortised
            lambda begin, count: net.grid(ctx, res, begin, count, self.sdf_scale, max(self.test_pointnum, 524288)),
            total, align=res * res)
        vals = grid[0].view(res, res, res).cpu().numpy()
        return vals, ctx

    def test(self, batch, eval_pred=False):
        vals, ctx = self.predict_grid(batch)
        mesh = generate_mesh(vals, self.bb_min, self.bb_max)
        scores = self.eval(mesh, batch.get("gt_mesh")) if eval_pred else {}
        occ = None
        return [mesh, occ, ctx.occ_pred.squeeze(1)], scores

    def eval(self, pred, gt):
        raise NotImplementedError("mesh quality metrics (reference evaluation/eval_util.py) are out of scope "
                                  "of the hot-path drop-in; evaluate the saved mesh with the reference's tools")

    def save(self, batch, pred, fname):
        mesh = pred[0]
        if mesh is None:
            raise RuntimeError("no mesh to save: PyMCubes/trimesh are not installed; use predict_grid()")
        mesh.export(fname + "_pred.obj")


def generate_mesh(gridvalues: np.ndarray, bb_min: float, bb_max: float):
    """utils.generate_mesh (utils.py:172-182): marching cubes of -grid at 0 through PyMCubes, when the
    optional host libraries are present; returns None otherwise (GPU marching cubes is a §8f row)."""
    try:
        import mcubes
        import trimesh
    except ImportError:
        return None
    vertices, triangles = mcubes.marching_cubes(-1.0 * gridvalues, 0)
    if len(vertices) > 10:
        vertices = (vertices - vertices.min()) / vertices.max()
        vertices = vertices * (bb_max - bb_min) + bb_min
    return trimesh.Trimesh(vertices, triangles)
