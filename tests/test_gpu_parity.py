"""GPU parity tests: the CUDA path through the C ABI against the oracle / the reference's golden
vectors.  Tolerances are BASELINE.json's: max|dSDF| <= 1e-4 in fp32, <= 2e-2 in bf16."""
import json
import os

import numpy as np
import pytest
import torch

from list_b200 import hotpath, synth
from oracle import list_oracle as O
from oracle import ref_port as P
from tests.helpers import GOLDEN, load_case, singular_mask

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4
BF16_TOL = 2e-2
DEV = "cuda:0"
CASES = ["small_camera_b2", "small_random_b1", "full_camera_b1", "full_random_b1", "train_b2"]


def ctx_and_weights(g, mode):
    ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, mode)
    return ctx, hotpath.prepare_weights(g.weights, ctx.layout, mode)


# ------------------------------------------------------------------ stage by stage (fp32)
def test_prep_maps_matches_upsample_oracle():
    inp = synth.make_inputs(seed=1, B=2, N=8, size="small")
    g = inp.to(DEV)
    ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "fp32")
    ref = O.prepare_maps(inp.maps)
    assert ctx.maps_cl.shape == ref.shape
    assert (ctx.maps_cl.cpu() - ref).abs().max().item() <= 2e-6
    ctx16 = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "bf16")
    assert (ctx16.maps_cl.float().cpu() - ref).abs().max().item() <= 2e-2
    assert torch.equal(ctx16.maps_cl.cpu(), ctx.maps_cl.bfloat16().cpu())


def test_prep_volume_is_an_exact_transpose():
    inp = synth.make_inputs(seed=2, B=2, N=8, size="small")
    g = inp.to(DEV)
    ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "fp32")
    for v, cl in zip(inp.vols, ctx.vols_cl):
        assert torch.equal(cl.cpu(), v.permute(0, 2, 3, 4, 1).contiguous())
    ctx16 = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "bf16")
    for v, cl in zip(inp.vols, ctx16.vols_cl):
        assert torch.equal(cl.cpu(), v.permute(0, 2, 3, 4, 1).contiguous().bfloat16())


@pytest.mark.parametrize("trans", ["camera", "random"])
def test_gather_rows_match_oracle_feature_rows(trans):
    inp = synth.make_inputs(seed=3, B=2, N=333, size="small", trans=trans)
    g = inp.to(DEV)
    ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "fp32")
    lay = ctx.layout
    X = hotpath.gather_features(ctx, g.points, raw=True).cpu()
    assert X.shape == (2 * 333, lay.k_pad)
    assert torch.equal(X[:, lay.k_out:], torch.zeros(2 * 333, lay.k_pad - lay.k_out))      # zero padding
    ref = O.feature_rows(inp.maps, inp.vols, inp.trans_mat, inp.points, fused_localise=True).reshape(-1, lay.k_out)
    got = torch.empty_like(ref)
    got[:, torch.from_numpy(lay.perm.astype(np.int64))] = X[:, :lay.k_out]                 # back to reference order
    bad = singular_mask(inp).reshape(-1)
    assert bad.float().mean() < 0.02
    assert (got - ref)[~bad].abs().max().item() <= 2e-5
    # swapped+scaled input convention gives the same rows
    q = (g.points[:, :, [2, 1, 0]] * 2).contiguous()
    assert torch.equal(hotpath.gather_features(ctx, q, raw=False).cpu(), X)


def test_gather_nan_divide_and_out_of_range_points():
    inp = synth.make_inputs(seed=4, B=1, N=64, size="small")
    inp.points = (torch.rand(1, 64, 3, generator=torch.Generator().manual_seed(5)) - 0.5) * 3.0   # beyond the box
    T = torch.zeros(1, 4, 3)
    T[0, 3, 2] = -1e-8            # h2 + 1e-8 == 0 and numerators 0 -> 0/0 (NaN grid -> zero 2-D features)
    inp.trans_mat = T
    g = inp.to(DEV)
    ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "fp32")
    lay = ctx.layout
    X = hotpath.gather_features(ctx, g.points).cpu()
    assert torch.isfinite(X).all()
    assert torch.equal(X[:, lay.map_off:lay.map_off + 1024], torch.zeros(64, 1024))
    ref = O.feature_rows(inp.maps, inp.vols, inp.trans_mat, inp.points).reshape(-1, lay.k_out)
    got = torch.empty_like(ref)
    got[:, torch.from_numpy(lay.perm.astype(np.int64))] = X[:, :lay.k_out]
    assert (got - ref).abs().max().item() <= 2e-5


def test_grid_points_bit_exact():
    z = np.load(os.path.join(GOLDEN, "grid_points.npz"))
    g5 = hotpath.grid_points(5).cpu().numpy()
    assert np.array_equal(g5, z["g5"].astype(np.float32))
    for res in (64, 256):
        ax = z[f"ax{res}_f32"]
        pts = hotpath.grid_points(res, begin=res * res * 3 + res * 5, count=res).cpu().numpy()   # x=3, y=5, z=0..res-1
        assert np.array_equal(pts[:, 2], ax) and np.all(pts[:, 0] == ax[3]) and np.all(pts[:, 1] == ax[5])
    full = hotpath.grid_points(64).cpu().numpy()
    assert np.array_equal(full[:130], z["g64_f32_head"]) and np.array_equal(full[-130:], z["g64_f32_tail"])
    assert np.array_equal(full, O.create_grid_points_from_bounds(-0.5, 0.5, 64).astype(np.float32))


def test_mlp_fp32_matches_oracle():
    gen = torch.Generator().manual_seed(6)
    lay = hotpath.feature_layout(1024, [1, 16, 32, 64, 128, 128])
    w = synth.mlp_weights(lay.k_out, gen)
    Xr = torch.randn(777, lay.k_out, generator=gen)                  # reference column order
    ref = O.implicit_mlp(Xr.unsqueeze(0), w)[0]
    X = torch.zeros(777, lay.k_pad)
    X[:, :lay.k_out] = Xr[:, torch.from_numpy(lay.perm.astype(np.int64))]
    kw = hotpath.prepare_weights({k: v.to(DEV) for k, v in w.items()}, lay, "fp32")
    out = hotpath.mlp(kw, X.to(DEV)).cpu()
    assert (out - ref).abs().max().item() <= 1e-5


# ------------------------------------------------------------------ end to end vs the reference's golden vectors
@pytest.mark.parametrize("name", CASES)
def test_fp32_path_vs_reference_golden(name):
    inp, z, _ = load_case(name)
    g = inp.to(DEV)
    ctx, kw = ctx_and_weights(g, "fp32")
    sdf = hotpath.query_sdf(ctx, kw, g.points, chunk_rows=1000).cpu().numpy()
    bad = singular_mask(inp).numpy()
    err = np.abs(sdf - z["sdf"])
    print(f"{name}: fp32 max|dSDF| = {err[~bad].max():.3e}, ill-conditioned points skipped: {int(bad.sum())}")
    assert bad.mean() < 0.01
    assert err[~bad].max() <= FP32_TOL


@pytest.mark.parametrize("name", CASES)
def test_bf16_path_vs_reference_golden(name):
    inp, z, _ = load_case(name)
    g = inp.to(DEV)
    ctx, kw = ctx_and_weights(g, "bf16")
    sdf = hotpath.query_sdf(ctx, kw, g.points).cpu().numpy()
    bad = singular_mask(inp, rel=1e-2).numpy()
    err = np.abs(sdf - z["sdf"])
    print(f"{name}: bf16 max|dSDF| = {err[~bad].max():.3e}")
    assert np.isfinite(sdf).all()
    assert err[~bad].max() <= BF16_TOL
    assert np.median(err) <= 2e-3          # far inside the bound on typical points


# ------------------------------------------------------------------ dense grid (a-8) and size-independent properties
def test_dense_grid_32_vs_oracle_and_sign_pattern():
    inp = synth.make_inputs(seed=8, B=1, N=8, size="small", trans="camera")
    ref = P.dense_grid_sdf(inp.maps, inp.vols, inp.trans_mat, inp.weights, 32, sdf_scale=10.0, chunk=8192)
    g = inp.to(DEV)
    ctx, kw = ctx_and_weights(g, "fp32")
    got = hotpath.grid_sdf(ctx, kw, 32, sdf_scale=10.0, chunk_rows=5000)[0].view(32, 32, 32).cpu().numpy()
    err = np.abs(got - ref)
    assert err.max() <= FP32_TOL
    # "identical marching-cubes topology": every vertex has the reference's sign, except vertices whose
    # |SDF| is within the fp32 tolerance of 0 (listed, must be inside the tolerance)
    flipped = np.sign(got) != np.sign(ref)
    assert np.all(np.abs(ref[flipped]) <= FP32_TOL)
    cells_changed = int((O.mc_case_index(got) != O.mc_case_index(ref)).sum())
    print(f"grid 32^3: max|dSDF|={err.max():.2e}, flipped vertices={int(flipped.sum())}, MC cells changed={cells_changed}")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("res", [24, 40])
def test_grid_kernel_vs_explicit_points_and_shards_compose(mode, res, monkeypatch):
    """The dense-grid walker kernel equals the per-point kernel on the same grid points up to fp32
    re-association (it evaluates the trilinear sum separably along z-runs), the generic kernel in grid
    mode equals it bit for bit, and a grid evaluated in shards / chunks of any size is bit-identical to
    one pass (what multi-GPU sharding relies on)."""
    inp = synth.make_inputs(seed=9, B=2, N=8, size="small", trans="random")
    g = inp.to(DEV)
    ctx, kw = ctx_and_weights(g, mode)
    total = res ** 3
    whole = hotpath.grid_sdf(ctx, kw, res, sdf_scale=10.0, chunk_rows=total)
    pts = hotpath.grid_points(res).unsqueeze(0).expand(2, -1, -1).contiguous()
    explicit = hotpath.query_sdf(ctx, kw, pts, out_div=10.0, chunk_rows=4096)
    # bf16: the dense-grid path hoists fc_0 through the samplers (csrc/hoist.cu), so it rounds to bf16 at different
    # places than the per-point path; both stay far inside the 2e-2 bound (values here are SDF / 10)
    tol = 2e-6 if mode == "fp32" else 2e-3
    assert (whole - explicit).abs().max().item() <= tol
    parts = []
    for begin, count in ((0, 1000), (1000, 5000), (6000, total - 6000)):
        parts.append(hotpath.grid_sdf(ctx, kw, res, begin, count, sdf_scale=10.0, chunk_rows=777))
    assert torch.equal(torch.cat(parts, dim=1), whole)
    monkeypatch.setenv("LIST_B200_GRID_GENERIC", "1")       # per-point kernel in grid mode ...
    monkeypatch.setenv("LIST_B200_HOIST", "0")              # ... on full (un-hoisted) feature rows
    generic = hotpath.grid_sdf(ctx, kw, res, sdf_scale=10.0, chunk_rows=3000)
    assert torch.equal(generic, explicit)


def test_grid_walker_rows_match_generic_rows():
    """Feature rows of the walker kernel against the per-point kernel, column by column."""
    inp = synth.make_inputs(seed=12, B=1, N=8, size="small", trans="camera")
    g = inp.to(DEV)
    for mode, tol in (("fp32", 5e-6), ("bf16", 0.04)):
        ctx, _ = ctx_and_weights(g, mode)
        for res, begin, count in ((20, 0, 8000), (33, 1234, 3000), (7, 0, 343)):
            a = hotpath.gather_grid_features(ctx, 0, res, begin, count).float()
            b = hotpath.gather_features(ctx, hotpath.grid_points(res, begin, count).unsqueeze(0)).float()
            assert a.shape == b.shape
            assert (a - b).abs().max().item() <= tol, (mode, res, (a - b).abs().max().item())


def test_point_permutation_and_batch_independence():
    inp = synth.make_inputs(seed=10, B=2, N=500, size="small")
    g = inp.to(DEV)
    ctx, kw = ctx_and_weights(g, "fp32")
    a = hotpath.query_sdf(ctx, kw, g.points)
    perm = torch.randperm(500, generator=torch.Generator().manual_seed(0)).to(DEV)
    b = hotpath.query_sdf(ctx, kw, g.points[:, perm].contiguous())
    assert torch.equal(a[:, perm], b)
    # image 1 alone gives the rows of image 1 in the batch
    one = synth.HotPathInputs([m[1:] for m in g.maps], [v[1:] for v in g.vols], g.trans_mat[1:], g.points[1:], g.weights)
    ctx1, _ = ctx_and_weights(one, "fp32")
    assert torch.equal(hotpath.query_sdf(ctx1, kw, one.points), a[1:])


def test_edge_sizes():
    inp = synth.make_inputs(seed=11, B=1, N=300, size="small")
    g = inp.to(DEV)
    for mode in ("fp32", "bf16"):
        ctx, kw = ctx_and_weights(g, mode)
        full = hotpath.query_sdf(ctx, kw, g.points)
        assert hotpath.query_sdf(ctx, kw, g.points[:, :0]).shape == (1, 0)
        for n in (1, 31, 33, 129, 257):
            assert torch.equal(hotpath.query_sdf(ctx, kw, g.points[:, :n].contiguous()), full[:, :n])


def test_full_size_256_grid_slab_properties():
    """BASELINE.json's full size (256^3 over the full-size feature tensors) through size-independent
    properties: a z-run of the grid evaluated by the grid kernel equals the explicit-point path, and
    fp32 and bf16 agree within the bf16 tolerance."""
    inp = synth.make_inputs(seed=333, B=1, N=8, size="full", trans="camera")
    g = inp.to(DEV)
    res = 256
    begin, count = (128 * res + 77) * res, 3 * res + 11           # crosses z-run boundaries
    out = {}
    for mode in ("fp32", "bf16"):
        ctx, kw = ctx_and_weights(g, mode)
        a = hotpath.grid_sdf(ctx, kw, res, begin, count, sdf_scale=10.0)
        pts = hotpath.grid_points(res, begin, count).unsqueeze(0)
        b = hotpath.query_sdf(ctx, kw, pts, out_div=10.0)
        assert (a - b).abs().max().item() <= (2e-6 if mode == "fp32" else 2e-3)
        out[mode] = a
    with torch.no_grad():
        ref = P.list_query(inp.maps, inp.vols, inp.trans_mat, hotpath.grid_points(res, begin, count).cpu().unsqueeze(0),
                           inp.weights) / 10.0
    assert (out["fp32"].cpu() - ref).abs().max().item() <= FP32_TOL
    assert (out["bf16"].cpu() - ref).abs().max().item() <= BF16_TOL


# ------------------------------------------------------------------ backward (a-9)
@pytest.mark.parametrize("glue", ["torch", "prep_fn"])
def test_backward_vs_reference_golden(glue):
    z = np.load(os.path.join(GOLDEN, "grad_small_b2.npz"))
    kw_ = json.loads(str(z["recipe"]))
    inp = synth.make_inputs(**kw_)
    _, sdf_gt = synth.training_points(kw_["B"], kw_["N"], torch.Generator().manual_seed(kw_["seed"] + 1000))
    g = inp.to(DEV)
    maps = [m.clone().requires_grad_(True) for m in g.maps]
    vols = [v.clone().requires_grad_(True) for v in g.vols]
    T = g.trans_mat.clone().requires_grad_(True)
    w = {k: v.clone().requires_grad_(True) for k, v in g.weights.items()}
    if glue == "torch":            # layouts built with stock differentiable torch ops
        ups = [torch.nn.functional.interpolate(m, size=137, mode="bilinear", align_corners=True) for m in maps]
        maps_cl = torch.cat(ups, dim=1).permute(0, 2, 3, 1).contiguous()
        vols_cl = [v.permute(0, 2, 3, 4, 1).contiguous() for v in vols]
    else:                          # what models.LIST.forward uses: the prep kernels as autograd functions
        maps_cl = hotpath.prep_maps_autograd(maps)
        vols_cl = [hotpath.prep_volume_autograd(v) for v in vols]
    sdf = hotpath.query_sdf_autograd(g.points, T, maps_cl, vols_cl, w, raw=True)
    assert np.abs(sdf.detach().cpu().numpy() - z["sdf"]).max() <= FP32_TOL
    loss = ((sdf_gt.to(DEV) * float(z["sdf_scale"]) - sdf) ** 2).sum(-1).mean()
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))

    def rel(a, b):
        a = a.detach().cpu().numpy()
        return np.abs(a - b).max() / (np.abs(b).max() + 1e-12)
    tol = 1e-3                                              # grads rel-err <= 1e-3 fp32
    assert rel(T.grad, z["dT"]) <= tol
    for key, p_ in w.items():                             # every MLP weight and bias
        gflat = p_.grad.flatten()
        if gflat.numel() > 70000:
            gflat = gflat[::97]                              # make_golden.py stores a strided subsample
        assert rel(gflat, z["dW_" + key.replace(".", "_")].reshape(-1)) <= tol, key
    for i, v in enumerate(vols):
        assert rel(v.grad.flatten()[::7], z[f"dvol{i}_sub"]) <= tol
        assert abs(v.grad.double().sum().item() - float(z[f"dvol{i}_sum"])) <= 1e-3 * (abs(float(z[f"dvol{i}_sum"])) + 1)
    for i, m in enumerate(maps):
        assert rel(m.grad.flatten()[::13], z[f"dmap{i}_sub"]) <= tol


# ------------------------------------------------------------------ hoisted fc_0 (dense grids, bf16)
def test_hoisted_rows_addend_and_verbatim_columns():
    """csrc/hoist.cu: the hoisted row is [addend(512) | the non-hoisted columns of the full row verbatim];
    the addend equals W0[:, :hoist_cols] applied to the hoisted columns of the full (walker) row, plus b0;
    the hoisted MLP (fc_0 on the remaining columns + addend in its epilogue) equals the MLP on full rows."""
    inp = synth.make_inputs(seed=21, B=2, N=8, size="small", trans="camera")
    g = inp.to(DEV)
    ctx, kw = ctx_and_weights(g, "bf16")
    ctx32, kw32 = ctx_and_weights(g, "fp32")
    hs = hotpath.HoistedState(ctx, kw)
    lay = ctx.layout
    hoist_cols = hs.hoist_cols
    assert hs.k_h == 512 + lay.k_pad - hoist_cols and hoist_cols == 1024 + 2 * 7 * 128
    for image in (0, 1):
        for res, begin, count in ((32, 0, 32 ** 3), (40, 12345, 20000), (33, 77, 3000), (64, 64 * 64 * 5 + 13, 500)):
            Xh_raw = hs.gather_grid(image, res, begin, count)
            Xh = Xh_raw.float()
            full_raw = hotpath.gather_grid_features(ctx, image, res, begin, count)
            assert torch.equal(Xh[:, 512:], full_raw.float()[:, hoist_cols:]), (res, begin)
            full32 = hotpath.gather_grid_features(ctx32, image, res, begin, count)
            want = full32[:, :hoist_cols] @ kw32.w0[:, :hoist_cols].t() + kw32.b0
            err = (Xh[:, :512] - want).abs().max().item()
            scale = want.abs().max().item()
            sdf_h = hs.mlp(Xh_raw)
            sdf_f = hotpath.mlp(kw, full_raw)
            d_sdf = (sdf_h - sdf_f).abs().max().item()
            print(f"image {image} res {res}: addend max err {err:.3e} (max |addend| {scale:.3f}), hoisted vs full MLP {d_sdf:.3e}")
            assert err <= 1e-2 * max(scale, 1.0)
            assert d_sdf <= 5e-3                      # both are bf16 evaluations of the same network output (O(0.1))
    # parts: the two gather kernels write disjoint column ranges
    a = torch.zeros(1000, hs.k_h, device=DEV, dtype=torch.bfloat16)
    hs.gather_grid(0, 32, 100, 1000, parts=1, out=a)
    assert torch.equal(a[:, 512:], torch.zeros_like(a[:, 512:])) and a[:, :512].abs().sum() > 0
    b = torch.zeros(1000, hs.k_h, device=DEV, dtype=torch.bfloat16)
    hs.gather_grid(0, 32, 100, 1000, parts=2, out=b)
    assert torch.equal(b[:, :512], torch.zeros_like(b[:, :512]))
    assert torch.equal(a + b, hs.gather_grid(0, 32, 100, 1000))


@pytest.mark.parametrize("res,begin,count", [(5, 0, 125), (5, 7, 60), (130, 130 * 130 * 64 + 77, 5000), (300, 300 * 300 * 150 + 299, 4000),
                                             (256, 256 * 256 * 200 + 100, 700)])
def test_hoisted_path_on_awkward_grid_sizes(res, begin, count, monkeypatch):
    """z-lines shorter than a tile, longer than two tiles, ranges starting / ending inside a z-line: the hoisted
    dense-grid path against the un-hoisted one on the same points (both bf16 evaluations of the same values)."""
    inp = synth.make_inputs(seed=23, B=1, N=8, size="small", trans="camera")
    g = inp.to(DEV)
    ctx, kw = ctx_and_weights(g, "bf16")
    a = hotpath.grid_sdf(ctx, kw, res, begin, count, sdf_scale=10.0, chunk_rows=1024)
    whole = hotpath.grid_sdf(ctx, kw, res, begin, count, sdf_scale=10.0, chunk_rows=count)
    assert torch.equal(a, whole)                             # chunking never changes a value
    monkeypatch.setenv("LIST_B200_HOIST", "0")
    b = hotpath.grid_sdf(ctx, kw, res, begin, count, sdf_scale=10.0, chunk_rows=1024)
    assert torch.isfinite(a).all() and (a - b).abs().max().item() <= 2e-3


def test_hoisted_grid_sdf_vs_oracle_and_unhoisted(monkeypatch):
    inp = synth.make_inputs(seed=22, B=1, N=8, size="small", trans="camera")
    ref = P.dense_grid_sdf(inp.maps, inp.vols, inp.trans_mat, inp.weights, 32, sdf_scale=10.0, chunk=8192).reshape(-1)
    g = inp.to(DEV)
    ctx, kw = ctx_and_weights(g, "bf16")
    hoisted = hotpath.grid_sdf(ctx, kw, 32, sdf_scale=10.0, chunk_rows=5000)[0].cpu().numpy()
    monkeypatch.setenv("LIST_B200_HOIST", "0")
    plain = hotpath.grid_sdf(ctx, kw, 32, sdf_scale=10.0, chunk_rows=5000)[0].cpu().numpy()
    e_h, e_p = np.abs(hoisted - ref).max(), np.abs(plain - ref).max()
    print(f"bf16 grid 32^3 vs oracle: hoisted {e_h:.3e}, unhoisted {e_p:.3e} (sdf/10)")
    assert e_h <= BF16_TOL and e_p <= BF16_TOL


def test_host_buffer_entry_points_equal_the_resident_path():
    """list_sdf_grid_host (C ABI, host pointers) and parallel.ShardedHostRunner (bench.py's e2e leg) against
    hotpath.grid_sdf on device-resident tensors: bit-identical SDF values."""
    from list_b200 import parallel
    inp = synth.make_inputs(seed=41, B=1, N=8, size="small", trans="camera")
    g = inp.to(DEV)
    res, begin, count = 48, 48 * 48 * 3, 48 * 48 * 20
    for mode in ("bf16", "fp32"):
        ctx, kw = ctx_and_weights(g, mode)
        want = hotpath.grid_sdf(ctx, kw, res, begin, count, sdf_scale=10.0, chunk_rows=8192)
        pin = lambda t: t.contiguous().pin_memory()
        maps, vols, T = [pin(m) for m in inp.maps], [pin(v) for v in inp.vols], pin(inp.trans_mat)
        c_path = hotpath.HostGridRunner(maps, vols, T, kw, res, begin, count, mode, 8192).run(10.0)
        torch.cuda.synchronize()
        assert torch.equal(c_path, want.cpu()), mode
        if mode == "bf16":
            runner = parallel.ShardedHostRunner(maps, vols, T, kw, res, mode, 8192)      # single rank: the whole grid
            out = runner.run(10.0)
            torch.cuda.synchronize()
            whole = hotpath.grid_sdf(ctx, kw, res, 0, res ** 3, sdf_scale=10.0, chunk_rows=8192)
            assert torch.equal(out, whole.cpu())


def test_host_buffer_call_overlaps_transfers_without_races():
    """list_sdf_grid_host uploads the big volumes while the projection / first addend gather run and downloads every
    chunk behind the next one's kernels.  Back-to-back calls on the same scratch with DIFFERENT inputs, two images, one
    and many chunks: each result must equal the resident path of its own inputs (a stale or half-uploaded tensor, or a
    download racing the next call's kernels, would show up here)."""
    res = 40
    pin = lambda t: t.contiguous().pin_memory()
    sets = [synth.make_inputs(seed=50 + i, B=2, N=8, size="small", trans="camera") for i in range(3)]
    for mode in ("bf16", "fp32"):
        for chunk in (4096, res ** 3):
            want = []
            _, kw = ctx_and_weights(sets[0].to(DEV), mode)             # one set of weights for every call
            for inp in sets:
                g = inp.to(DEV)
                ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, mode)
                want.append(hotpath.grid_sdf(ctx, kw, res, 0, res ** 3, sdf_scale=2.0, chunk_rows=chunk).cpu())
            torch.cuda.synchronize()
            # one runner: same pinned host tensors and device scratch, refilled between calls
            maps, vols, T = [pin(m) for m in sets[0].maps], [pin(v) for v in sets[0].vols], pin(sets[0].trans_mat)
            runner = hotpath.HostGridRunner(maps, vols, T, kw, res, 0, res ** 3, mode, chunk)
            for i, inp in enumerate(sets):
                for dst, src in zip([*maps, *vols, T], [*inp.maps, *inp.vols, inp.trans_mat]):
                    dst.copy_(src)
                got = runner.run(2.0)                      # no sync between enqueue and the next host-side refill ...
                torch.cuda.synchronize()                   # ... except this one: the host tensors are reused
                assert torch.equal(got, want[i]), (mode, chunk, i)
            # two calls enqueued back to back without a host sync in between (same inputs): the second must not disturb
            # the first one's downloads, and both leave the same values
            runner.run(2.0)
            got = runner.run(2.0)
            torch.cuda.synchronize()
            assert torch.equal(got, want[-1]), (mode, chunk, "back-to-back")


def test_staged_grid_call_prepares_late_levels_itself():
    """list_sdf_grid_late (the multi-rank e2e leg): levels that are still being produced on another stream when the
    call is made -- here: copied from pinned host memory behind a long sleep kernel -- are waited for and prepared by
    the call; the context's buffers of those levels hold NaN until then, so a premature read cannot go unnoticed.  The
    pinned host copy of the result arrives chunk by chunk.  Bit-identical to the plain call."""
    inp = synth.make_inputs(seed=61, B=2, N=8, size="small", trans="camera")
    g = inp.to(DEV)
    res, begin, count = 40, 40 * 40 * 2, 40 * 40 * 30
    late_levels = [l for l, v in enumerate(inp.vols) if v.shape[1] % 64 != 0]          # never read by the projection
    assert late_levels
    side = torch.cuda.Stream(device=DEV)
    for mode in ("bf16", "fp32"):
        ctx, kw = ctx_and_weights(g, mode)
        want = hotpath.grid_sdf(ctx, kw, res, begin, count, sdf_scale=3.0, chunk_rows=8192).cpu()
        for chunk in (8192, count):
            ctx2 = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, mode, skip_levels=late_levels)
            raw = [None] * len(inp.vols)
            pinned = {l: inp.vols[l].contiguous().pin_memory() for l in late_levels}
            for l in late_levels:
                ctx2.vols_cl[l].fill_(float("nan"))
                raw[l] = torch.full_like(g.vols[l], float("nan"))
            ws = hotpath._workspace(ctx2.struct(), kw.struct(), min(chunk, count), DEV)
            out = torch.empty(2, count, device=DEV, dtype=torch.float32)
            out_host = torch.zeros(2, count, dtype=torch.float32).pin_memory()
            ev = torch.cuda.Event()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                torch.cuda._sleep(200_000_000)                                          # ~0.1 s: the call is enqueued long before
                for l in late_levels:
                    raw[l].copy_(pinned[l], non_blocking=True)
                ev.record(side)
            hotpath.grid_sdf_late(ctx2, kw, res, begin, count, 3.0, chunk, out, ws, ev, raw, out_host=out_host)
            torch.cuda.synchronize()
            assert torch.equal(out.cpu(), want), (mode, chunk)
            assert torch.equal(out_host, want), (mode, chunk, "host copy")
    # a level the projection reads cannot be late
    ctx, kw = ctx_and_weights(g, "bf16")
    hoisted = [l for l, v in enumerate(inp.vols) if v.shape[1] % 64 == 0]
    raw = [g.vols[l] if l == hoisted[-1] else None for l in range(len(inp.vols))]
    ws = hotpath._workspace(ctx.struct(), kw.struct(), 8192, DEV)
    with pytest.raises(RuntimeError, match="cannot be late"):
        hotpath.grid_sdf_late(ctx, kw, res, 0, 8192, 1.0, 8192, torch.empty(2, 8192, device=DEV), ws, None, raw)


# ------------------------------------------------------------------ SURVEY.md §8d parity gates at the configured sizes
def _camera_inputs():
    return synth.make_inputs(seed=synth.SEED, B=1, N=8, size="full", trans="camera")


def test_parity_gate_cfg1_full_64_cubed_grid():
    """cfg-1: every point of the 64^3 grid over the full-size per-image tensors, against the ATen-op oracle on the
    CPU: fp32 <= 1e-4 with the same sign at every vertex (up to |sdf| <= 1e-4), bf16 <= 2e-2 (network units)."""
    inp = _camera_inputs()
    ref = P.dense_grid_sdf(inp.maps, inp.vols, inp.trans_mat, inp.weights, 64, sdf_scale=1.0)          # (64,64,64)
    g = inp.to(DEV)
    out = {}
    for mode in ("fp32", "bf16"):
        ctx, kw = ctx_and_weights(g, mode)
        out[mode] = hotpath.grid_sdf(ctx, kw, 64, sdf_scale=1.0, chunk_rows=65536)[0].view(64, 64, 64).cpu().numpy()
    e32, e16 = np.abs(out["fp32"] - ref), np.abs(out["bf16"] - ref)
    flipped = np.sign(out["fp32"]) != np.sign(ref)
    print(f"cfg-1 64^3: fp32 max|dSDF| {e32.max():.2e}, bf16 max {e16.max():.2e} (median {np.median(e16):.1e}), "
          f"flipped vertices {int(flipped.sum())}, MC cells changed {int((O.mc_case_index(out['fp32']) != O.mc_case_index(ref)).sum())}")
    assert e32.max() <= FP32_TOL and e16.max() <= BF16_TOL
    assert np.all(np.abs(ref[flipped]) <= FP32_TOL)


def test_parity_gate_cfg4_strided_subsample_of_256_cubed():
    """cfg-4: the whole 256^3 grid is evaluated on the GPU (bf16 production path and fp32 parity path); every 64th
    point (262 144 of them) is checked against the oracle on the CPU."""
    inp = _camera_inputs()
    res, stride = 256, 64
    g = inp.to(DEV)
    idx = torch.arange(5, res ** 3, stride)
    pts = hotpath.grid_points(res).cpu()[idx].unsqueeze(0)
    with torch.no_grad():
        ref = torch.cat([P.list_query(inp.maps, inp.vols, inp.trans_mat, p, inp.weights) for p in torch.split(pts, 65536, 1)], 1)[0] / 10.0
    for mode, tol in (("bf16", BF16_TOL / 10.0), ("fp32", FP32_TOL / 10.0)):          # values are SDF / sdf_scale
        ctx, kw = ctx_and_weights(g, mode)
        full = hotpath.grid_sdf(ctx, kw, res, sdf_scale=10.0, chunk_rows=1048576 if mode == "bf16" else 131072)[0]
        err = (full.cpu()[idx] - ref).abs()
        print(f"cfg-4 256^3 subsample, {mode}: max|dSDF|/scale {err.max().item():.2e} over {idx.numel()} points")
        assert err.max().item() <= tol


def test_parity_gate_cfg3_128_cubed_grid_and_mesh_topology():
    """cfg-3: the whole 128^3 grid on the GPU in both precisions; every 64th point against the oracle, and -- the
    north_star's "identical marching-cubes topology on the fp32 path" -- the marching-cubes case index of every cell of a
    dense 4-plane slab of the grid (65 536 vertices evaluated by the oracle) on the fp32 grid."""
    inp = _camera_inputs()
    res, stride = 128, 64
    g = inp.to(DEV)
    idx = torch.arange(3, res ** 3, stride)
    pts = hotpath.grid_points(res).cpu()
    with torch.no_grad():
        ref = P.list_query(inp.maps, inp.vols, inp.trans_mat, pts[idx].unsqueeze(0), inp.weights)[0] / 10.0
        x0 = 61
        slab = slice(x0 * res * res, (x0 + 4) * res * res)
        ref_slab = (P.list_query(inp.maps, inp.vols, inp.trans_mat, pts[slab].unsqueeze(0), inp.weights)[0] / 10.0).view(4, res, res).numpy()
    for mode, tol in (("bf16", BF16_TOL / 10.0), ("fp32", FP32_TOL / 10.0)):
        ctx, kw = ctx_and_weights(g, mode)
        full = hotpath.grid_sdf(ctx, kw, res, sdf_scale=10.0, chunk_rows=524288 if mode == "bf16" else 131072)[0].cpu()
        err = (full[idx] - ref).abs().max().item()
        got_slab = full[slab].view(4, res, res).numpy()
        changed = int((O.mc_case_index(got_slab) != O.mc_case_index(ref_slab)).sum())
        print(f"cfg-3 128^3, {mode}: max|dSDF|/scale {err:.2e} over {idx.numel()} points; marching-cubes cells changed in the slab: {changed}")
        assert err <= tol
        if mode == "fp32":
            flipped = np.sign(got_slab) != np.sign(ref_slab)
            assert np.all(np.abs(ref_slab[flipped]) <= FP32_TOL / 10.0)
            assert changed == int(flipped.any()) * changed              # cells may only change through a listed sign flip
            assert changed == 0 or flipped.sum() > 0


def test_parity_gate_cfg5_eight_images_per_gpu():
    """cfg-5: one call over B = 8 full-size images (the per-GPU share of 64 images on 8 B200), 128^3 each, bf16 line-table
    path; 4 096 strided points of every image against the oracle."""
    B, res, stride = 8, 128, 512
    inp = synth.make_inputs(seed=synth.SEED + 5, B=B, N=8, size="full", trans="camera")
    g = inp.to(DEV)
    ctx, kw = ctx_and_weights(g, "bf16")
    full = hotpath.grid_sdf(ctx, kw, res, sdf_scale=10.0, chunk_rows=1048576).cpu()
    assert full.shape == (B, res ** 3)
    idx = torch.arange(7, res ** 3, stride)
    pts = hotpath.grid_points(res).cpu()[idx].unsqueeze(0).expand(B, -1, -1).contiguous()
    with torch.no_grad():
        ref = P.list_query(inp.maps, inp.vols, inp.trans_mat, pts, inp.weights) / 10.0
    err = (full[:, idx] - ref).abs()
    print(f"cfg-5 8 x 128^3 bf16: max|dSDF|/scale per image {[f'{e:.1e}' for e in err.max(dim=1).values.tolist()]}")
    assert err.max().item() <= BF16_TOL / 10.0
    # image b of the batch equals the same image evaluated alone (no cross-image state in the hoisted tensors)
    one = synth.HotPathInputs([m[3:4] for m in g.maps], [v[3:4] for v in g.vols], g.trans_mat[3:4], g.points[3:4], g.weights)
    ctx1, _ = ctx_and_weights(one, "bf16")
    alone = hotpath.grid_sdf(ctx1, kw, res, begin=res ** 3 // 2, count=65536, sdf_scale=10.0, chunk_rows=65536).cpu()
    assert torch.equal(alone[0], full[3, res ** 3 // 2: res ** 3 // 2 + 65536])


def test_prep_adjoints_match_torch_autograd():
    """list_prep_maps_bwd / list_prep_volume_bwd (row a-9: autograd of reference modules.py:25-35 and of the layout change)
    against torch's own backward of F.interpolate(align_corners=True) + permute, on full-size tensors."""
    import torch.nn.functional as F
    inp = synth.make_inputs(seed=91, B=2, N=8, size="full", trans="camera")
    g = inp.to("cuda:0")
    maps = [m.clone().requires_grad_(True) for m in g.maps]
    vols = [v.clone().requires_grad_(True) for v in g.vols]
    maps_cl = hotpath.prep_maps_autograd(maps)
    vols_cl = [hotpath.prep_volume_autograd(v) for v in vols]
    gen = torch.Generator(device="cuda:0").manual_seed(5)
    up = torch.randn(maps_cl.shape, device="cuda:0", generator=gen)
    ups = [torch.randn(v.shape, device="cuda:0", generator=gen) for v in vols_cl]
    torch.autograd.backward([maps_cl, *vols_cl], [up, *ups])
    ours = [m.grad.clone() for m in maps] + [v.grad.clone() for v in vols]
    maps2 = [m.detach().clone().requires_grad_(True) for m in g.maps]
    vols2 = [v.detach().clone().requires_grad_(True) for v in g.vols]
    ref_cl = torch.cat([F.interpolate(m, size=(137, 137), mode="bilinear", align_corners=True) for m in maps2], dim=1).permute(0, 2, 3, 1)
    assert torch.allclose(ref_cl, maps_cl, atol=1e-5, rtol=1e-5)
    refv = [v.permute(0, 2, 3, 4, 1) for v in vols2]
    torch.autograd.backward([ref_cl, *refv], [up, *ups])
    theirs = [m.grad for m in maps2] + [v.grad for v in vols2]
    for i, (a, b) in enumerate(zip(ours, theirs)):
        assert a.shape == b.shape and a.is_contiguous()
        err = (a - b).abs().max().item()
        assert err <= 1e-4 * max(1.0, b.abs().max().item()), (i, err)
