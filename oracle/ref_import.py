"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference from /root/reference (this container only; the
GPU box has no /root/reference) with the four shims SURVEY.md §8c lists:
 1. empty stub modules for `mcubes` / `trimesh` (utils.py:3,7 import them; the
    hot path never calls them);
 2. torchvision.models.resnet18(pretrained=True) -> weights=None (no network);
 3. torch.Tensor.cuda -> identity on a GPU-less host (modules.py:214 calls
    .cuda() in a constructor; forward re-homes the table, modules.py:256);
 4. a plain config object instead of arguments.get_args().
No reference source is copied or edited.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("LIST_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "network"))


class RefConfig:
    """Fields LIST.__init__ reads (models.py:39-71; arguments.py defaults)."""
    vox_res = 128
    im_enc_layers = [1, 1, 1, 1, 16, 32, 64, 128, 128]
    train_batch_size = 1
    point_feat = [128, 128, 256, 256, 256, 128, 128, 3]
    point_degree = [2, 2, 2, 2, 2, 2, 64]
    bb_min = -0.5
    bb_max = 0.5


_loaded = None


def load():
    """Returns (network.modules, network.models) of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for name in ("mcubes", "trimesh"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    import torchvision.models as tvm
    if not getattr(tvm.resnet18, "_list_shim", False):
        _orig = tvm.resnet18

        def resnet18(pretrained=False, **kw):
            kw.pop("weights", None)
            return _orig(weights=None, **kw)
        resnet18._list_shim = True
        tvm.resnet18 = resnet18
    if not torch.cuda.is_available() and not getattr(torch.Tensor.cuda, "_list_shim", False):
        def _cuda(self, *a, **k):
            return self
        _cuda._list_shim = True
        torch.Tensor.cuda = _cuda
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    cwd = os.getcwd()
    try:
        os.chdir(REFERENCE_ROOT)          # modules.py:12 sets a relative TORCH_HOME
        import network.modules as ref_modules
        import network.models as ref_models
    finally:
        os.chdir(cwd)
    _loaded = (ref_modules, ref_models)
    return _loaded
