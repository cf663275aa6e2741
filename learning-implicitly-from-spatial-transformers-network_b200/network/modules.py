"""Mirror of the reference's `network.modules` for the LIST model.

Hot path (CUDA, via the C ABI -- no PyTorch arithmetic):
  * PerceptualPooling  (reference network/modules.py:15-59)
  * VoxelDecoder2      (reference network/modules.py:192-214, 247-282)
Per-image stages (stock PyTorch, run once per image; out of the hot path) are
re-stated only so that `LIST` exists with the reference's attribute names and state_dict keys:
  * ResEncoder (modules.py:1027-1074), PointMLP (:62-104), TreeGraphDecoder (:107-132 +
    layers/gcn.py:6-69), VoxelEncoder2 (:401-442).
"""
from __future__ import annotations

import math
import os
import warnings
import weakref
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import hotpath

DISPLACEMENT = 0.0722     # reference modules.py:205


# =============================================================================== hot path
def _inference_only(name: str, *tensors) -> None:
    """The stand-alone hot modules run the kernels on detached tensors.  Under autograd that would silently train
    nothing (reference call pattern models.py:93-97), so it is an error: training goes through `LIST.forward`, where
    gather + MLP are ONE autograd node (hotpath.query_sdf_autograd)."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise RuntimeError(
            f"list_b200 {name}.forward is inference-only (its inputs require grad): call it under torch.no_grad(), or "
            "train through list_b200.network.models.LIST.forward, whose hot path is differentiable")


def _cached(cache: dict, tensors, build):
    """`build()` once per set of input tensor OBJECTS (weak references + version counters: a new tensor that happens to
    reuse the storage address of a freed one never hits), per device."""
    slot = str(tensors[0].device)
    hit = cache.get(slot)
    if hit is not None and len(hit[0]) == len(tensors) and all(r() is t and v == t._version for (r, v), t in zip(hit[0], tensors)):
        return hit[1]
    value = build()
    cache[slot] = ([(weakref.ref(t), t._version) for t in tensors], value)
    return value


class PerceptualPooling(nn.Module):
    """forward(img_featuremaps: 5 x (B,C_i,H_i,W_i), pc: (B,N,3), trans_mat: (B,4,3)) -> (B, sum C_i, 1, N)

    Same signature and result as the reference module.  Stand-alone use (the reference's own
    call pattern, executors.py:219-220) runs the gather kernel on a maps-only context and returns
    the 2-D feature block transposed to the reference's layout; inside `LIST` the fused path
    (`LIST.query`) is used instead so the features never leave the row matrix.
    """

    def __init__(self, map_size: int = 137):
        super().__init__()
        self.map_size = map_size
        self._ctx_cache = {}

    def forward(self, img_featuremaps: Sequence[torch.Tensor], pc: torch.Tensor, trans_mat: torch.Tensor) -> torch.Tensor:
        _inference_only("PerceptualPooling", *img_featuremaps, pc, trans_mat)
        B, N, _ = pc.shape
        # the reference re-runs the 137^2 upsample of all five maps for every 65 536-point chunk (modules.py:25-35 inside
        # executors.py:215-224); here the prepared context is kept while the caller passes the same tensors again
        ctx = _cached(self._ctx_cache, (*img_featuremaps, trans_mat), lambda: hotpath.prepare_context(
            img_featuremaps, [torch.zeros(B, 8, 1, 1, 1, device=pc.device)], trans_mat, "fp32", self.map_size))
        X = hotpath.gather_features(ctx, pc, raw=False)
        lay = ctx.layout
        cm = ctx.maps_cl.shape[-1]
        feats = X[:, lay.map_off:lay.map_off + cm].reshape(B, N, cm)
        return feats.permute(0, 2, 1).unsqueeze(2).contiguous()

    def __repr__(self):
        return f"{self.__class__.__name__} (Map pc to {self.map_size} x {self.map_size} plane)"


class VoxelDecoder2(nn.Module):
    """forward(p: (B,N,3), feat: 6 x (B,C,D,H,W), percep_feat: (B,1024,N)) -> (B,N)

    Parameters live under the reference's names (`fc.fc_0.weight` [2*h_dim, feature_size, 1], ...,
    modules.py:196-200) so checkpoints load unchanged; kernel-format copies are derived on demand
    and cached against the parameters' version counters.
    Stand-alone use with an explicit `percep_feat` follows the reference's data flow: the voxel
    part comes from the gather kernel, the given perceptual block is placed in the row matrix and
    the MLP kernel runs on it.
    """

    def __init__(self, feature_size: int, h_dim: int):
        super().__init__()
        self.feature_size = feature_size
        self.fc = nn.ModuleDict()
        self.fc["fc_0"] = nn.Conv1d(feature_size, h_dim * 2, 1)
        self.fc["fc_1"] = nn.Conv1d(h_dim * 2, h_dim, 1)
        self.fc["fc_2"] = nn.Conv1d(h_dim, h_dim, 1)
        self.fc["fc_out"] = nn.Conv1d(h_dim, 1, 1)
        self.actvn = nn.ReLU()
        rows = [[0.0, 0.0, 0.0]]
        for axis in range(3):
            for sign in (-1, 1):
                r = [0.0, 0.0, 0.0]
                r[axis] = sign * DISPLACEMENT
                rows.append(r)
        # a plain attribute like the reference's (not a buffer -> not in the state_dict); unlike the
        # reference it is NOT moved with .cuda() at construction, so the module builds on CPU hosts
        self.displacments = torch.tensor(rows)
        self._cache = {}
        self._ctx_cache = {}

    # ---- derived kernel weights -------------------------------------------------------------
    def param_dict(self, prefix: str = "fc.") -> dict:
        return {f"{prefix}{k}.{n}": getattr(m, n) for k, m in self.fc.items() for n in ("weight", "bias")}

    def kernel_weights(self, layout, dtype="fp32") -> "hotpath.KernelWeights":
        """Kernel-format copies of the parameters, rebuilt when a parameter changes (version counter / storage).  The cache
        is keyed per (device, dtype) and the LOCALLY built object is returned: nn.DataParallel replicas share this dict
        (shallow __dict__ copy, reference train.py:126) and call from parallel threads, one device each."""
        params = self.param_dict()
        dev = next(iter(params.values())).device
        slot = (str(dev), hotpath.dtype_code(dtype))
        key = (layout.k_pad, tuple(layout.perm[:8]), tuple((p.data_ptr(), p._version) for p in params.values()))
        hit = self._cache.get(slot)
        if hit is not None and hit[0] == key:
            return hit[1]
        kw = hotpath.prepare_weights(params, layout, dtype)
        self._cache[slot] = (key, kw)
        return kw

    def forward(self, p: torch.Tensor, feat: Sequence[torch.Tensor], percep_feat: torch.Tensor) -> torch.Tensor:
        _inference_only("VoxelDecoder2", p, percep_feat, *feat, *self.parameters())
        B, N, _ = p.shape
        cm = percep_feat.shape[1]
        ctx = _cached(self._ctx_cache, tuple(feat), lambda: hotpath.prepare_context(
            [torch.zeros(B, cm, 2, 2, device=p.device)], feat, torch.zeros(B, 4, 3, device=p.device), "fp32", 2))
        X = hotpath.gather_features(ctx, p, raw=False)
        lay = ctx.layout
        X[:, lay.map_off:lay.map_off + cm] = percep_feat.detach().permute(0, 2, 1).reshape(B * N, cm)
        sdf = hotpath.mlp(self.kernel_weights(lay, "fp32"), X)
        return sdf.view(B, N)


# =============================================================================== per-image stages
class PointMLP(nn.Module):
    """3 -> 64 -> 256 -> 512 pointwise conv + BN + ReLU on the coarse cloud (modules.py:62-104)."""

    def __init__(self):
        super().__init__()

        def block(i, o):
            return nn.Sequential(nn.Conv2d(i, o, 1, 1), nn.BatchNorm2d(o), nn.ReLU(inplace=True))
        self.block1, self.block2, self.block3 = block(3, 64), block(64, 256), block(256, 512)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_normal_(m.weight)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):                       # (B,N,3) -> (B,512,1,N)
        x = x.transpose(1, 2).unsqueeze(2)
        return self.block3(self.block2(self.block1(x)))


class TreeGCN(nn.Module):
    """One tree-structured graph-conv layer (layers/gcn.py:6-69): ancestors term + branching + loop."""

    def __init__(self, batch, depth, features, degrees, support=10, node=1, upsample=False, activation=True):
        super().__init__()
        self.depth, self.node, self.degree = depth, node, degrees[depth]
        self.in_feature, self.out_feature = features[depth], features[depth + 1]
        self.upsample, self.activation = upsample, activation
        self.W_root = nn.ModuleList([nn.Linear(features[i], self.out_feature, bias=False) for i in range(depth + 1)])
        if upsample:
            self.W_branch = nn.Parameter(torch.empty(node, self.in_feature, self.degree * self.in_feature))
        self.W_loop = nn.Sequential(nn.Linear(self.in_feature, self.in_feature * support, bias=False),
                                    nn.Linear(self.in_feature * support, self.out_feature, bias=False))
        self.bias = nn.Parameter(torch.empty(1, self.degree, self.out_feature))
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2)
        if upsample:
            nn.init.kaiming_normal_(self.W_branch.data, a=0.2, mode="fan_in", nonlinearity="leaky_relu")
        bound = 1.0 / math.sqrt(self.out_feature)
        self.bias.data.uniform_(-bound, bound)

    def forward(self, tree: List[torch.Tensor]):
        B = tree[-1].size(0)
        root = 0
        for i in range(self.depth + 1):
            rep = self.node // tree[i].size(1)
            root = root + self.W_root[i](tree[i]).repeat(1, 1, rep).view(B, -1, self.out_feature)
        if self.upsample:
            br = self.leaky_relu(tree[-1].unsqueeze(2) @ self.W_branch)
            br = self.W_loop(br.view(B, self.node * self.degree, self.in_feature))
            out = root.repeat(1, 1, self.degree).view(B, -1, self.out_feature) + br
        else:
            out = root + self.W_loop(tree[-1])
        if self.activation:
            out = self.leaky_relu(out + self.bias.repeat(1, self.node, 1))
        tree.append(out)
        return tree


class TreeGraphDecoder(nn.Module):
    """Global code -> coarse point cloud (modules.py:107-132)."""

    def __init__(self, batch_size, features, degrees, support):
        super().__init__()
        assert len(features) - 1 == len(degrees), "Number of features should be one more than number of degrees."
        self.batch_size, self.layer_num = batch_size, len(degrees)
        self.gcn = nn.Sequential()
        nodes = 1
        for i in range(self.layer_num):
            self.gcn.add_module(f"TreeGCN_{i}", TreeGCN(batch_size, i, features, degrees, support=support, node=nodes,
                                                        upsample=True, activation=(i != self.layer_num - 1)))
            nodes *= degrees[i]

    def forward(self, tree):
        return self.gcn(tree)[-1]


class ResEncoder(nn.Module):
    """ResNet-18 with a stride-1 stem -> (128-d code, 5 feature maps) (modules.py:1027-1074).
    Like the reference (`models.resnet18(pretrained=True)`, modules.py:1030) the backbone starts from the ImageNet weights
    when torchvision has them in its local cache; without them (this image has no network) it warns once and starts from
    random init -- loading a LIST checkpoint overwrites every backbone tensor either way.  `pretrained=False` skips the
    lookup."""
    _warned = False

    def __init__(self, pretrained: Optional[bool] = None):
        super().__init__()
        from torchvision import models
        weights = None
        if pretrained is None or pretrained:
            w = models.ResNet18_Weights.DEFAULT
            cached = os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(w.url))
            if os.path.exists(cached):
                weights = w
            elif not ResEncoder._warned:
                ResEncoder._warned = True
                warnings.warn(f"ResEncoder: ImageNet weights not found at {cached}; starting from random init "
                              "(the reference starts from torchvision's pretrained resnet18)")
        net = models.resnet18(weights=weights)
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=1, padding=3, bias=False)
        self.bn1, self.relu, self.maxpool = net.bn1, net.relu, net.maxpool
        self.layer1, self.layer2, self.layer3, self.layer4 = net.layer1, net.layer2, net.layer3, net.layer4
        self.avgpool, self.fc = net.avgpool, net.fc
        self.fc1 = nn.Linear(1000, 128)

    def forward(self, x):
        f0 = self.relu(self.bn1(self.conv1(x)))
        f1 = self.layer1(self.maxpool(f0))
        f2 = self.layer2(f1)
        f3 = self.layer3(f2)
        f4 = self.layer4(f3)
        code = self.fc1(self.fc(torch.flatten(self.avgpool(f4), 1)))
        return code, [f0, f1, f2, f3, f4]


class VoxelEncoder2(nn.Module):
    """Occupancy grid -> 6-level feature pyramid (modules.py:401-442)."""

    def __init__(self, layers):
        super().__init__()
        self.layers = layers
        self.conv = nn.ModuleDict()
        self.bn = nn.ModuleList()
        self.relu, self.sigmoid, self.maxpool = nn.ReLU(), nn.Sigmoid(), nn.MaxPool3d(2)
        for l in range(len(layers) - 1):
            self.conv[f"conv_{l}"] = nn.Conv3d(layers[l], layers[l + 1], 3, padding=1)
            if l > 2:
                self.conv[f"conv_{l}_0"] = nn.Conv3d(layers[l + 1], layers[l + 1], 3, padding=1)
            self.bn.append(nn.BatchNorm3d(layers[l + 1]))

    def forward(self, x):
        feats = []
        net = x.unsqueeze(1)
        for l in range(len(self.layers) - 1):
            if l < 2:
                net = self.bn[l](self.relu(self.conv[f"conv_{l}"](net)))
            elif l == 2:
                net = self.sigmoid(self.conv[f"conv_{l}"](net))
                feats.append(net)
            else:
                net = self.relu(self.conv[f"conv_{l}"](net))
                net = self.bn[l](self.relu(self.conv[f"conv_{l}_0"](net)))
                feats.append(net)
                net = self.maxpool(net)
        return feats
