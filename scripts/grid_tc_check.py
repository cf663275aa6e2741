"""Stage-by-stage check of the line-table dense-grid path on a B200 (csrc/lines.cu, csrc/grid_tc.cu): prints every
comparison instead of asserting, so one GPU call shows where a discrepancy starts.

    python scripts/grid_tc_check.py [--case NAME]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from list_b200 import hotpath, synth                 # noqa: E402
from oracle import list_oracle as O                  # noqa: E402
from oracle import ref_port                          # noqa: E402

DISP = O.DISPLACEMENT


def axis(c, R):
    i = (((c + 1.0) / 2.0) * (R - 1)).clamp(0.0, float(R - 1))
    f = i.floor()
    i0 = f.long()
    return i0, (i0 + 1).clamp(max=R - 1), (f + 1.0) - i, i - f


def table_reference(ctx, kw, lay, ls, res, lines, dev):
    """fp32 emulation of G for the given z-lines from the bf16 volumes and bf16 W0 blocks (GPU torch)."""
    ax = torch.tensor(np.linspace(-0.5, 0.5, res), dtype=torch.float32, device=dev) * 2
    disp = O.displacements().to(dev)
    out = torch.zeros(len(lines), ls.rows_per_line, 512, device=dev)
    hoisted = [l for l in range(len(ctx.vols_cl) - 1, -1, -1) if ctx.vol_ch[l] % 8 == 0 and lay.vol_off[l] < ls.hoist_cols]
    rowbase = 0
    lz = torch.tensor([ln // res for ln in lines], device=dev)
    ly = torch.tensor([ln % res for ln in lines], device=dev)
    qy, qz = ax[ly], ax[lz]
    for l in hoisted:
        V = ctx.vols_cl[l][0].float()
        R, Cc = V.shape[0], V.shape[3]
        for cls, dl in ((0, [0, 3, 4, 5, 6]), (1, [1]), (2, [2])):
            acc = torch.zeros(len(lines), R, 512, device=dev)
            for d in dl:
                Wd = kw.w0[:, lay.vol_off[l] + d * Cc: lay.vol_off[l] + (d + 1) * Cc].float()      # (512, C)
                PV = (V.reshape(-1, Cc) @ Wd.t()).to(torch.bfloat16).float().reshape(R, R, R, 512)
                z0, z1, wz0, wz1 = axis(qz + disp[d, 2], R)
                y0, y1, wy0, wy1 = axis(qy + disp[d, 1], R)
                for zi, wz in ((z0, wz0), (z1, wz1)):
                    for yi, wy in ((y0, wy0), (y1, wy1)):
                        acc += (wy * wz).view(-1, 1, 1) * PV[zi, yi]
            out[:, rowbase + cls * R: rowbase + (cls + 1) * R] = acc
        rowbase += 3 * R
    return out


def run_case(name, res, begin, count, trans, dev, size="full"):
    print(f"==== case {name}: res {res} begin {begin} count {count} T={trans} tensors={size}", flush=True)
    inp = synth.make_inputs(seed=synth.SEED, B=1, N=8, size=size, trans=trans)
    g = inp.to(dev)
    ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "bf16")
    kw = hotpath.prepare_weights(g.weights, ctx.layout, "bf16")
    lay = ctx.layout
    grid = torch.tensor(O.create_grid_points_from_bounds(-0.5, 0.5, res)[begin:begin + count]).unsqueeze(0).float()
    t0 = time.time()
    with torch.no_grad():
        ref = ref_port.list_query(inp.maps, inp.vols, inp.trans_mat, grid, inp.weights)[0]
    print(f"oracle: {time.time() - t0:.1f} s, |sdf| max {ref.abs().max():.3f}")

    ls = hotpath.LineTableState(ctx, kw)
    print(f"layout: hoist_cols {ls.hoist_cols} k_f {ls.k_f} rows/line {ls.rows_per_line} hoist buffer {ls.buf.numel() / 2**20:.1f} MiB")
    G = ls.table(0, res, begin, count)
    torch.cuda.synchronize()
    line0 = begin // res
    nl = G.shape[0]
    pick = sorted(set([0, nl - 1, nl // 2] + list(np.random.default_rng(0).integers(0, nl, size=min(nl, 6)))))
    Gref = table_reference(ctx, kw, lay, ls, res, [line0 + i for i in pick], dev)
    dG = (G[pick].float() - Gref).abs()
    print(f"G table: {nl} lines; max|dG| {dG.max().item():.3e} vs max|G| {Gref.abs().max().item():.3e} (bf16 rounding expected ~4e-3 relative)")

    Xr = ls.rest(0, res, begin, count)
    Xfull = hotpath.gather_grid_features(ctx, 0, res, begin, count)
    torch.cuda.synchronize()
    dX = (Xr.float() - Xfull[:, ls.hoist_cols:ls.hoist_cols + ls.k_f].float()).abs().max().item()
    print(f"Xr vs the full gather's columns [{ls.hoist_cols}, +{ls.k_f}): max|d| {dX:.3e} (expected 0)")

    plan = ls.plan(0, res, begin, count, G)
    sdf, h1, tr = ls.evaluate(res, begin, count, Xr, plan, 1.0, debug=True, trace=True)
    torch.cuda.synchronize()
    h1_ref = torch.relu(Xfull.float() @ kw.w0.float().t() + kw.b0)
    dh = (h1 - h1_ref).abs()
    print(f"relu(fc_0): max|d| {dh.max().item():.3e} mean {dh.mean().item():.3e} vs max {h1_ref.max().item():.3f}; "
          f"rows with max|d| > 0.05: {(dh.max(dim=1).values > 0.05).sum().item()} of {count}")
    if dh.max().item() > 0.05:
        bad = torch.nonzero(dh.max(dim=1).values > 0.05).flatten()[:16].tolist()
        print("  first bad rows (row, step on line):", [(r, (begin + r) % res) for r in bad])
        r = bad[0]
        print("  row", r, "h1[:8]", h1[r, :8].tolist(), "ref", h1_ref[r, :8].tolist())
    err = (sdf.cpu() - ref).abs()
    print(f"sdf vs oracle: max {err.max().item():.3e} median {err.median().item():.3e} (tolerance 2e-2)")
    sdf_full = hotpath.mlp(kw, Xfull)
    print(f"sdf vs unhoisted bf16 MLP on the full rows: max {(sdf - sdf_full).abs().max().item():.3e}")
    try:
        hs = hotpath.HoistedState(ctx, kw)
        sdf_old = hs.mlp(hs.gather_grid(0, res, begin, count))
        print(f"sdf vs round-1 hoisted path: max {(sdf - sdf_old).abs().max().item():.3e}")
    except RuntimeError as e:
        print("round-1 hoisted path unavailable:", e)
    # the default entry point, two chunkings
    a = hotpath.grid_sdf(ctx, kw, res, begin, count, 1.0, chunk_rows=count)
    b = hotpath.grid_sdf(ctx, kw, res, begin, count, 1.0, chunk_rows=max(1, count // 3 + 17))
    torch.cuda.synchronize()
    print(f"list_sdf_grid: one chunk vs stages max|d| {(a[0] - sdf).abs().max().item():.3e}; three chunks vs one: "
          f"{(a - b).abs().max().item():.3e} (expected 0 and 0)")
    t = tr.cpu().numpy()
    if t[0, 0] != 0:
        d = lambda i, j: float(np.median(t[2:12, j] - t[2:12, i]))
        print("trace (cycles, median of tiles 2..11): fc_0 issue", d(0, 1), "fc0 done->ep0 done", d(6, 7), "tile", float(np.median(t[3:12, 0] - t[2:11, 0])))
    return err.max().item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="all")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cases = {
        "small32": (32, 0, 32 ** 3, "camera", "small"),
        "r128": (128, 128 * 128 * 37 + 128 * 5, 128 * 64, "camera", "full"),
        "r128mid": (128, 128 * 128 * 37 + 128 * 5 + 37, 128 * 40 + 11, "camera", "full"),
        "r256": (256, 256 * 256 * 100 + 256 * 31, 256 * 32, "camera", "full"),
        "r256rand": (256, 256 * 256 * 128 + 256 * 100, 256 * 32, "random", "full"),
        "r64": (64, 64 * 64 * 20, 64 * 96, "camera", "full"),
    }
    worst = 0.0
    for name, (res, begin, count, trans, size) in cases.items():
        if a.case not in ("all", name):
            continue
        worst = max(worst, run_case(name, res, begin, count, trans, dev, size))
    print("worst sdf error", worst)


if __name__ == "__main__":
    main()
