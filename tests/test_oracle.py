"""CPU tests: the oracle (explicit restatement + ATen port) against the golden vectors the
UNMODIFIED reference produced (oracle/make_golden.py), plus the reference itself when
/root/reference is present (this container)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import list_oracle as O
from oracle import ref_port as P
from oracle import ref_import
from tests.helpers import GOLDEN, load_case, singular_mask

FWD_CASES = ["small_camera_b2", "small_random_b1", "full_camera_b1", "full_random_b1", "train_b2"]
FP32_TOL = 1e-4        # BASELINE.json north_star: max |dSDF| <= 1e-4 in fp32


@pytest.mark.parametrize("name", FWD_CASES)
def test_port_matches_reference_golden(name):
    inp, z, _ = load_case(name)
    with torch.no_grad():
        sdf = P.list_query(inp.maps, inp.vols, inp.trans_mat, inp.points, inp.weights)
    assert sdf.shape == z["sdf"].shape
    # same ATen ops on the same torch build: equal up to thread-count dependent blocking
    assert np.abs(sdf.numpy() - z["sdf"]).max() <= 2e-6


@pytest.mark.parametrize("name", FWD_CASES)
def test_restatement_matches_reference_golden(name):
    if name.startswith("full") and os.environ.get("LIST_FAST_TESTS"):
        pytest.skip("fast mode")
    inp, z, _ = load_case(name)
    with torch.no_grad():
        sdf = O.list_query(inp.maps, inp.vols, inp.trans_mat, inp.points, inp.weights)
    bad = singular_mask(inp).numpy()
    err = np.abs(sdf.numpy() - z["sdf"])
    assert bad.mean() < 0.01
    assert err[~bad].max() <= FP32_TOL / 4, err[~bad].max()


def test_restatement_percep_features():
    inp, z, _ = load_case("small_camera_b2")
    q = inp.points[:, :, [2, 1, 0]] * 2
    xy, _ = O.localise(q, inp.trans_mat)
    f = O.gather2d(O.prepare_maps(inp.maps), xy)           # (B,N,1024)
    ref = torch.from_numpy(z["percep_head"])               # (B,1024,16)
    assert (f[:, :16].transpose(1, 2) - ref).abs().max() <= 2e-5


def test_upsample_restatement_vs_aten():
    g = torch.Generator().manual_seed(1)
    for size_in in (3, 14, 28, 56, 224):
        x = torch.randn(2, 5, size_in, size_in, generator=g)
        a = O.upsample_bilinear_align_corners(x, 137)
        b = torch.nn.functional.interpolate(x, size=137, mode="bilinear", align_corners=True)
        assert (a - b).abs().max() <= 2e-6


def test_gather3d_restatement_vs_aten_border_cases():
    g = torch.Generator().manual_seed(2)
    vol = torch.randn(1, 4, 5, 6, 7, generator=g)
    pts = torch.cat([torch.rand(1, 200, 3, generator=g) * 2.6 - 1.3,          # in and out of range
                     torch.tensor([[[1.0, 1.0, 1.0], [-1.0, -1.0, -1.0], [1.0, -1.0, 0.0], [0.0, 0.0, 0.0]]])], 1)
    a = O.gather3d(O.to_channels_last_vol(vol), pts)
    b = torch.nn.functional.grid_sample(vol, pts.view(1, 1, 1, -1, 3), padding_mode="border", align_corners=True)
    assert (a - b.view(4, -1).t().unsqueeze(0)).abs().max() <= 2e-6


def test_gather2d_nan_and_edges():
    g = torch.Generator().manual_seed(3)
    m = torch.randn(1, 137, 137, 8, generator=g)
    xy = torch.tensor([[[0.0, 0.0], [136.0, 136.0], [136.0, 0.5], [float("nan"), 3.0], [67.99999, 68.00001]]])
    a = O.gather2d(m, xy)
    grid = ((xy - 68.0) / 68.0).unsqueeze(1)
    b = torch.nn.functional.grid_sample(m.permute(0, 3, 1, 2).contiguous(), grid, align_corners=True)
    # NaN grid (0/0 in the divide): ATen's CUDA sampler maps non-finite coordinates out of bounds
    # (zeros); its vectorised CPU sampler returns NaN.  The restatement follows the CUDA
    # behaviour, which is what the reference's GPU runs see.
    assert torch.equal(a[0, 3], torch.zeros(8))
    keep = [0, 1, 2, 4]
    assert (a[:, keep] - b.view(8, -1).t().unsqueeze(0)[:, keep]).abs().max() <= 2e-6


def test_displacement_table_order():
    d = O.displacements()
    assert d.shape == (7, 3)
    assert torch.equal(d[0], torch.zeros(3))
    exp = [(0, -1), (0, 1), (1, -1), (1, 1), (2, -1), (2, 1)]
    for row, (ax, s) in zip(d[1:], exp):
        assert abs(row[ax].item() - s * 0.0722) < 1e-7 and row.abs().sum().item() == pytest.approx(0.0722, rel=1e-6)


def test_grid_points_golden():
    z = np.load(os.path.join(GOLDEN, "grid_points.npz"))
    assert np.array_equal(O.create_grid_points_from_bounds(-0.5, 0.5, 5), z["g5"])
    g64 = O.create_grid_points_from_bounds(-0.5, 0.5, 64).astype(np.float32)
    assert np.array_equal(g64[:130], z["g64_f32_head"]) and np.array_equal(g64[-130:], z["g64_f32_tail"])
    assert np.allclose(g64.astype(np.float64).sum(0), z["g64_sum"], atol=1e-9)


def test_backward_restatement_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN, "grad_small_b2.npz"))
    kw = json.loads(str(z["recipe"]))
    from list_b200 import synth
    inp = synth.make_inputs(**kw)
    _, sdf_gt = synth.training_points(kw["B"], kw["N"], torch.Generator().manual_seed(kw["seed"] + 1000))
    maps = [m.clone().requires_grad_(True) for m in inp.maps]
    vols = [v.clone().requires_grad_(True) for v in inp.vols]
    T = inp.trans_mat.clone().requires_grad_(True)
    w = {k: v.clone().requires_grad_(True) for k, v in inp.weights.items()}
    sdf = O.list_query(maps, vols, T, inp.points, w)
    loss = O.sdf_loss(sdf, sdf_gt, float(z["sdf_scale"]))
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))

    def rel(a, b):
        return np.abs(a - b).max() / (np.abs(b).max() + 1e-12)
    assert rel(T.grad.numpy(), z["dT"]) <= 1e-3
    assert rel(w["fc.fc_out.weight"].grad.numpy(), z["dW_fc_fc_out_weight"]) <= 1e-4
    assert rel(w["fc.fc_2.weight"].grad.numpy(), z["dW_fc_fc_2_weight"]) <= 1e-4
    assert rel(w["fc.fc_0.weight"].grad.flatten()[::97].numpy(), z["dW_fc_fc_0_weight"]) <= 1e-4
    for i, v in enumerate(vols):
        assert rel(v.grad.flatten()[::7].numpy(), z[f"dvol{i}_sub"]) <= 1e-4
    for i, m in enumerate(maps):
        assert rel(m.grad.flatten()[::13].numpy(), z[f"dmap{i}_sub"]) <= 1e-4


def test_mc_case_index_sign_equivalence():
    rng = np.random.default_rng(0)
    g = rng.standard_normal((6, 6, 6)).astype(np.float32)
    a = O.mc_case_index(g)
    b = O.mc_case_index(g * 3.0)
    assert np.array_equal(a, b) and a.shape == (5, 5, 5)
    g2 = g.copy()
    g2[2, 2, 2] = -g2[2, 2, 2]
    assert (O.mc_case_index(g2) != a).sum() == 8


@pytest.mark.skipif(not ref_import.available(), reason="needs /root/reference (build container only)")
def test_port_and_restatement_vs_live_reference():
    from list_b200 import synth
    from oracle.make_golden import ref_forward
    ref_modules, _ = ref_import.load()
    inp = synth.make_inputs(seed=4242, B=2, N=300, size="small", trans="camera")
    with torch.no_grad():
        ref = ref_forward(ref_modules, inp)
        port = P.list_query(inp.maps, inp.vols, inp.trans_mat, inp.points, inp.weights)
        rest = O.list_query(inp.maps, inp.vols, inp.trans_mat, inp.points, inp.weights)
    assert (port - ref).abs().max() <= 2e-6
    assert (rest - ref).abs().max() <= FP32_TOL / 4


def test_projection_commutes_with_the_samplers():
    """The algebra behind the hoisted fc_0 of the dense-grid path (csrc/hoist.cu), checked on the
    oracle alone: fc_0 is linear and so are the bilinear (zeros padding) and trilinear (border padding) samplers, so
    projecting a feature tensor through its block of W0 and sampling the 512-wide result equals sampling first and
    multiplying after -- including taps that fall outside the map, NaN grids, clamped voxel coordinates and the seven
    displaced copies; and the unshifted trilinear weights sum to one, which is what lets a constant (the bias) ride
    along in a projected volume."""
    g = torch.Generator().manual_seed(5)
    B, N, S, C, R, Cv, n0 = 2, 257, 9, 16, 5, 8, 12
    maps_cl = torch.randn(B, S, S, C, generator=g, dtype=torch.float64)
    vol_cl = torch.randn(B, R, R, R, Cv, generator=g, dtype=torch.float64)
    Wm = torch.randn(n0, C, generator=g, dtype=torch.float64)
    Wv = torch.randn(7, n0, Cv, generator=g, dtype=torch.float64)                 # one W0 block per displacement
    xy = (torch.rand(B, N, 2, generator=g, dtype=torch.float64) * 1.4 - 0.2) * (S - 1)   # some taps out of bounds
    xy[0, 0] = float("nan")
    xy[0, 1] = torch.tensor([S - 1.0, S - 1.0])
    q = torch.rand(B, N, 3, generator=g, dtype=torch.float64) * 2.6 - 1.3        # beyond the border on all sides
    # a-3: sample then project == project then sample
    a = O.gather2d(maps_cl, xy) @ Wm.t()
    b = O.gather2d(maps_cl @ Wm.t(), xy)
    assert torch.allclose(a, b, rtol=0, atol=1e-12)
    # a-5: per displacement d, its own W0 block
    disp = O.displacements(torch.float64)
    lhs = sum(O.gather3d(vol_cl, q + disp[d]) @ Wv[d].t() for d in range(7))
    rhs = sum(O.gather3d(vol_cl @ Wv[d].t(), q + disp[d]) for d in range(7))
    assert torch.allclose(lhs, rhs, rtol=0, atol=1e-11)
    # a constant rides along exactly once per sample: trilinear weights sum to one under border padding
    ones = torch.ones(B, R, R, R, 1, dtype=torch.float64)
    assert torch.allclose(O.gather3d(ones, q), torch.ones(B, N, 1, dtype=torch.float64), rtol=0, atol=1e-13)
    # ... but not the bilinear weights with zeros padding (why the bias is not folded into the projected map)
    ones2 = torch.ones(B, S, S, 1, dtype=torch.float64)
    w = O.gather2d(ones2, xy)
    assert (w < 1 - 1e-6).any() and torch.all(w <= 1 + 1e-12)
