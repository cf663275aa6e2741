"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol
include/list_b200.h declares, and its host-only entry points behave (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from list_b200 import _C, hotpath

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "list_b200.h")).read()
    declared = set(re.findall(r"LIST_API\s+[\w\s\*]+?\b(list_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    assert declared == set(_C.SIGNATURES), declared ^ set(_C.SIGNATURES)
    lib = _C.lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.list_b200_abi_version() == _C.ABI_VERSION


def test_struct_sizes_match_the_header_layout():
    assert C.sizeof(_C.ListLayout) == 4 * 4 + 8 * 4
    assert C.sizeof(_C.ListCtx) == 16 + 8 + 8 + 32 + 32 + 64 + 8
    assert C.sizeof(_C.ListWeights) == 24 + 8 * 8
    assert C.sizeof(_C.ListGrads) == 8 + 64 + 8 + 8 * 8


def test_feature_layout_default_config():
    lay = hotpath.feature_layout(1024, [1, 16, 32, 64, 128, 128])
    assert (lay.k_out, lay.k_pad) == (3610, 3648)          # reference models.py:43
    assert sorted(lay.perm.tolist()) == list(range(3610))  # a permutation of the reference columns
    assert lay.map_off % 8 == 0 and all(o % 8 == 0 for o, c in zip(lay.vol_off, [1, 16, 32, 64, 128, 128]) if c % 8 == 0)
    # reference column of (level, c, d) is (cum_c + c)*7 + d; percep at 2583.., q at 3607..
    assert lay.perm[lay.vol_off[0] + 3] == 3                # level 0 (C=1), d=3
    assert lay.perm[lay.vol_off[1] + 2 * 16 + 5] == (1 + 5) * 7 + 2
    assert lay.perm[lay.map_off + 10] == 2583 + 10
    assert lay.perm[lay.xyz_off + 2] == 3609


def test_feature_layout_other_channel_configs():
    lay = hotpath.feature_layout(64, [8, 3])
    assert lay.k_out == 64 + 7 * 11 + 3 and lay.k_pad % 64 == 0
    assert sorted(lay.perm.tolist()) == list(range(lay.k_out))


def test_error_convention_bad_arguments():
    lib = _C.lib()
    lay = _C.ListLayout()
    rc = lib.list_feature_layout(1023, 1, _C.i32_array([8]), C.byref(lay), None)
    assert rc == _C.EINVAL and "multiple of 8" in _C.last_error()
    rc = lib.list_grid_points(None, 4, -0.5, 0.5, 0, 100, None)
    assert rc == _C.EINVAL
    assert lib.list_mlp_workspace_bytes(None, 10) == 0


def test_cpu_tensors_are_rejected_loudly():
    import torch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hotpath.prepare_context([torch.zeros(1, 8, 4, 4)], [torch.zeros(1, 8, 2, 2, 2)], torch.zeros(1, 4, 3))
    with pytest.raises(RuntimeError):
        hotpath.grid_points(4, device="cpu")


def test_staged_and_host_entry_points_reject_bad_arguments_without_a_gpu():
    """Argument validation of the host-buffer / staged dense-grid calls happens before any CUDA work."""
    lib = _C.lib()
    rc = lib.list_sdf_grid_late(None, None, 8, -0.5, 0.5, 0, 8, None, 1.0, 8, None, 0, None, None, None, None)
    assert rc == _C.EINVAL and "ctx is NULL" in _C.last_error()
    rc = lib.list_sdf_grid_host(None, None, None, 5, 137, None, 6, None, None, None, 1, _C.BF16, None, 8, -0.5, 0.5, 0, 8,
                                1.0, 8, None, None, 0, None)
    assert rc == _C.EINVAL and "NULL argument" in _C.last_error()
    assert lib.list_sdf_grid_host_bytes(None, None, 5, 137, 6, None, None, 1, _C.BF16, 8, 8) == 0


def test_round2_entry_points_reject_bad_arguments_without_a_gpu():
    """The entry points added in round 2 validate before any CUDA work, and the size queries are pure host arithmetic."""
    lib = _C.lib()
    assert lib.list_sdf_grid_workspace_bytes(None, None, 64, 1024) == 0
    rc = lib.list_prep_maps_bwd(None, None, None, 5, 1, 137, None, None)
    assert rc == _C.EINVAL and "NULL argument" in _C.last_error()
    rc = lib.list_prep_volume_bwd(None, 1, 16, 8, None, None)
    assert rc == _C.EINVAL and "NULL argument" in _C.last_error()
    rc = lib.list_grid_tc_fwd(None, None, None, 8, -0.5, 0.5, 0, 8, None, 384, None, None, 1.0, None, None, None, None)
    assert rc == _C.EINVAL
    # plans: 19 KB per tile of 128 steps (+ header); a 256^3 chunk of 4 M rows has 32768 tiles
    n = lib.list_grid_plan_bytes(256, 0, 4194304)
    assert 32768 * (832 * 8 + 128 * 24 * 4) <= n <= 32768 * (832 * 8 + 128 * 24 * 4) + 32768 * 4 + 512
    assert lib.list_grid_plan_bytes(256, 0, 0) == 0
