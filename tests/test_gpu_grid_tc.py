"""GPU tests of the line-table dense-grid path (csrc/lines.cu, csrc/grid_tc.cu), the default of list_sdf_grid in bf16
mode: per-line column tables, tile plans, and the fused interpolation + MLP kernel, stage by stage and end to end
against the ATen-op oracle.  Tolerance: BASELINE.json's 2e-2 for the bf16 tensor-core mode (scaled SDF)."""
import numpy as np
import pytest
import torch

from list_b200 import hotpath, synth
from oracle import list_oracle as O
from oracle import ref_port as P

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2
DEV = "cuda:0"


def _setup(seed, size, trans, B=1):
    inp = synth.make_inputs(seed=seed, B=B, N=8, size=size, trans=trans)
    g = inp.to(DEV)
    ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "bf16")
    kw = hotpath.prepare_weights(g.weights, ctx.layout, "bf16")
    return inp, g, ctx, kw


def _axis(c, R):
    i = (((c + 1.0) / 2.0) * (R - 1)).clamp(0.0, float(R - 1))
    f = i.floor()
    i0 = f.long()
    return i0, (i0 + 1).clamp(max=R - 1), (f + 1.0) - i, i - f


def _table_reference(ctx, kw, ls, image, res, lines):
    """fp32 emulation of the line tables from the bf16 volumes and bf16 W0 blocks (reference modules.py:205-212, 262-265:
    displacement table, trilinear / border / align_corners), projected volumes rounded to bf16 as hoist.cu stores them."""
    lay = ctx.layout
    ax = torch.tensor(np.linspace(-0.5, 0.5, res), dtype=torch.float32, device=DEV) * 2
    disp = O.displacements().to(DEV)
    out = torch.zeros(len(lines), ls.rows_per_line, 512, device=DEV)
    hoisted = [l for l in range(len(ctx.vols_cl) - 1, -1, -1) if ctx.vol_ch[l] % 8 == 0 and lay.vol_off[l] < ls.hoist_cols]
    lz = torch.tensor([ln // res for ln in lines], device=DEV)
    ly = torch.tensor([ln % res for ln in lines], device=DEV)
    qy, qz = ax[ly], ax[lz]
    rowbase = 0
    for l in hoisted:
        V = ctx.vols_cl[l][image].float()
        R, Cc = V.shape[0], V.shape[3]
        for cls, dl in ((0, [0, 3, 4, 5, 6]), (1, [1]), (2, [2])):
            acc = torch.zeros(len(lines), R, 512, device=DEV)
            for d in dl:
                Wd = kw.w0[:, lay.vol_off[l] + d * Cc: lay.vol_off[l] + (d + 1) * Cc].float()
                PV = (V.reshape(-1, Cc) @ Wd.t()).to(torch.bfloat16).float().reshape(R, R, R, 512)
                z0, z1, wz0, wz1 = _axis(qz + disp[d, 2], R)
                y0, y1, wy0, wy1 = _axis(qy + disp[d, 1], R)
                for zi, wz in ((z0, wz0), (z1, wz1)):
                    for yi, wy in ((y0, wy0), (y1, wy1)):
                        acc += (wy * wz).view(-1, 1, 1) * PV[zi, yi]
            out[:, rowbase + cls * R: rowbase + (cls + 1) * R] = acc
        rowbase += 3 * R
    return out


@pytest.mark.parametrize("size,res,begin,count", [("small", 24, 0, 24 ** 3), ("small", 40, 12345, 20000), ("full", 128, 128 * 128 * 37 + 128 * 5 + 37, 5000)])
def test_line_tables_match_the_emulation(size, res, begin, count):
    inp, g, ctx, kw = _setup(31, size, "camera", B=2 if size == "small" else 1)
    ls = hotpath.LineTableState(ctx, kw)
    assert ls.k_f == ctx.layout.k_pad - ls.hoist_cols and ls.rows_per_line > 0
    for image in range(ctx.B):
        G = ls.table(image, res, begin, count)
        nl = (begin + count - 1) // res - begin // res + 1
        assert G.shape == (nl, ls.rows_per_line, 512)
        pick = sorted(set([0, nl - 1, nl // 2, nl // 3]))
        ref = _table_reference(ctx, kw, ls, image, res, [begin // res + i for i in pick])
        err = (G[pick].float() - ref).abs().max().item()
        scale = ref.abs().max().item()
        print(f"{size} res {res} image {image}: max|dG| {err:.3e} of {scale:.3f}")
        assert err <= 1.2e-2 * max(scale, 1.0)            # two bf16 roundings of O(1) values


CASES = [
    ("small", 32, 0, 32 ** 3, "camera"),
    ("small", 33, 77, 3000, "random"),                               # z-lines that are not a multiple of anything
    ("full", 64, 64 * 64 * 20, 64 * 96, "camera"),
    ("full", 128, 128 * 128 * 37 + 128 * 5 + 37, 128 * 40 + 11, "camera"),   # starts and ends inside a z-line
    ("full", 256, 256 * 256 * 100 + 256 * 31, 256 * 32, "camera"),
    ("full", 256, 256 * 256 * 128 + 256 * 100, 256 * 32, "random"),  # clamped queries, the divide's singular plane
    ("full", 300, 300 * 300 * 150 + 299, 4000, "camera"),            # three tiles per z-line, the last one short
]


@pytest.mark.parametrize("size,res,begin,count,trans", CASES)
def test_stages_against_the_oracle_and_the_unhoisted_rows(size, res, begin, count, trans):
    """Xr equals the full gather's non-hoisted columns bit for bit; relu(fc_0) of the fused kernel (dense part on Xr +
    interpolated part on the tensor cores) equals fc_0 on the full feature rows up to bf16 rounding; the SDF meets the
    bf16 tolerance against the fp32 oracle; list_sdf_grid (any chunking) returns exactly the stage results."""
    inp, g, ctx, kw = _setup(synth.SEED, size, trans)
    grid = torch.tensor(O.create_grid_points_from_bounds(-0.5, 0.5, res)[begin:begin + count]).unsqueeze(0).float()
    with torch.no_grad():
        ref = P.list_query(inp.maps, inp.vols, inp.trans_mat, grid, inp.weights)[0]
    ls = hotpath.LineTableState(ctx, kw)
    G = ls.table(0, res, begin, count)
    Xr = ls.rest(0, res, begin, count)
    Xfull = hotpath.gather_grid_features(ctx, 0, res, begin, count)
    expect = Xfull[:, ls.hoist_cols:ls.hoist_cols + ls.k_f].clone()
    bias_col = ctx.layout.xyz_off + 3 - ls.hoist_cols                  # fc_0's bias rides in the MMA: 1.0 in three pad columns
    expect[:, bias_col:bias_col + 3] = 1.0
    assert torch.equal(Xr, expect)
    plan = ls.plan(0, res, begin, count, G)
    stats = torch.zeros(2, device=DEV, dtype=torch.int64)
    sdf, h1 = ls.evaluate(res, begin, count, Xr, plan, 1.0, debug=True, stats=stats)
    h1_ref = torch.relu(Xfull.float() @ kw.w0.float().t() + kw.b0)
    dh = (h1 - h1_ref).abs().max().item()
    err = (sdf.cpu() - ref).abs().max().item()
    pairs, ksteps = (int(x) for x in stats.cpu())
    chunks = ksteps / 4
    print(f"{size} res {res} T={trans}: relu(fc_0) max|d| {dh:.3e}, sdf max|d| {err:.3e}, {chunks / max(pairs, 1):.2f} interpolation chunks per tile pair")
    assert torch.isfinite(sdf).all()
    assert dh <= 3e-2 and err <= BF16_TOL
    whole = hotpath.grid_sdf(ctx, kw, res, begin, count, 1.0, chunk_rows=count)
    assert torch.equal(whole[0], sdf)
    parts = hotpath.grid_sdf(ctx, kw, res, begin, count, 1.0, chunk_rows=max(1, count // 3 + 17))
    assert torch.equal(parts, whole)


def test_line_table_path_equals_the_addend_path_within_bf16(monkeypatch):
    """Both hoisted formulations (round 1: addend block through HBM; now: interpolation inside the MLP kernel) are bf16
    evaluations of the same values; two images per call."""
    inp, g, ctx, kw = _setup(23, "small", "camera", B=2)
    res, begin, count = 40, 12345, 20000
    a = hotpath.grid_sdf(ctx, kw, res, begin, count, sdf_scale=10.0, chunk_rows=4096)
    monkeypatch.setenv("LIST_B200_LINES", "0")
    b = hotpath.grid_sdf(ctx, kw, res, begin, count, sdf_scale=10.0, chunk_rows=4096)
    monkeypatch.setenv("LIST_B200_HOIST", "0")
    c = hotpath.grid_sdf(ctx, kw, res, begin, count, sdf_scale=10.0, chunk_rows=4096)
    assert torch.isfinite(a).all()
    assert (a - b).abs().max().item() <= 2e-3 and (a - c).abs().max().item() <= 2e-3


def test_zero_weights_give_the_bias_chain():
    """Known answer: with W0 = 0 the fused kernel must return fc_out(relu(fc_2(relu(fc_1(relu(b0)))))) for every point --
    the interpolation chunks multiply rows of all-zero projected tensors, padding rows included."""
    inp = synth.make_inputs(seed=5, B=1, N=8, size="small", trans="random")
    inp.weights["fc.fc_0.weight"].zero_()
    g = inp.to(DEV)
    ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "bf16")
    kw = hotpath.prepare_weights(g.weights, ctx.layout, "bf16")
    got = hotpath.grid_sdf(ctx, kw, 24, sdf_scale=1.0)[0]
    w = {k: v.to(DEV) for k, v in inp.weights.items()}
    bf = lambda x: x.to(torch.bfloat16).float()
    h = bf(torch.relu(w["fc.fc_0.bias"]))
    h = bf(torch.relu(bf(w["fc.fc_1.weight"].squeeze(-1)) @ h + w["fc.fc_1.bias"]))
    h = torch.relu(bf(w["fc.fc_2.weight"].squeeze(-1)) @ h + w["fc.fc_2.bias"])
    want = (w["fc.fc_out.weight"].squeeze(-1) @ h + w["fc.fc_out.bias"]).item()
    assert (got - want).abs().max().item() <= 1e-5


@pytest.mark.parametrize("size,res,begin,count", [("small", 40, 12345, 20000), ("full", 128, 128 * 128 * 37 + 128 * 5 + 37, 70000),
                                                  ("full", 256, 256 * 256 * 100 + 256 * 31, 256 * 300)])
def test_tensor_core_line_tables_match_the_simt_kernel(monkeypatch, size, res, begin, count):
    """lines_tc.cu (opt-in, LIST_B200_LINES_TC=1): the H interpolation of the tables as a GEMM on the tensor cores with TMA
    tensor stores, against lines.cu's SIMT kernel.  One more bf16 rounding (the D-reduced rows) plus bf16 H weights."""
    inp, g, ctx, kw = _setup(37, size, "camera")
    ls = hotpath.LineTableState(ctx, kw)
    monkeypatch.delenv("LIST_B200_LINES_TC", raising=False)
    G0 = ls.table(0, res, begin, count)
    monkeypatch.setenv("LIST_B200_LINES_TC", "1")
    G1 = ls.table(0, res, begin, count)
    torch.cuda.synchronize()
    scale = G0.float().abs().max().item()
    err = (G1.float() - G0.float()).abs().max().item()
    print(f"{size} res {res}: max|dG| {err:.3e} of {scale:.3f}")
    assert err <= 1.2e-2 * max(scale, 1.0)
