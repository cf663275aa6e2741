"""CPU emulation of the tensor-core interpolation scheme (csrc/grid_tc.cu) to bound its error before any CUDA exists.

The hoisted fc_0 contribution of the maps and of the coarse levels is evaluated as  A_w[steps x rows] . B[rows x 512]
on the tensor cores: B rows are bf16 rows of the projected map P (pixel taps) and of the per-line column table G
((H,D)-interpolated, class-summed projected volumes), A_w holds the bf16-rounded interpolation weights.  This script
emulates every rounding of that path (bf16 operands, fp32 accumulation) on a few z-lines of the 256^3 grid with
full-size tensors and compares the SDF with the fp32 oracle.

    python scripts/tc_interp_prototype.py [--lines 24] [--res 256] [--trans camera|random]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from list_b200 import synth                      # noqa: E402
from oracle import list_oracle as O              # noqa: E402

DISP = O.DISPLACEMENT


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def axis(c, R):
    i = (((c + 1.0) / 2.0) * (R - 1)).clamp(0.0, float(R - 1))
    f = i.floor()
    i0 = f.long()
    i1 = (i0 + 1).clamp(max=R - 1)
    return i0, i1, (f + 1.0) - i, i - f


def quant_pair(w0, w1, mode):
    """bf16 weights of a two-tap lerp.  'unit': round the larger, complement the smaller (sum is exactly 1)."""
    if mode == "exact":
        return w0, w1
    if mode == "rn":
        return bf(w0), bf(w1)
    big0 = w0 >= w1
    a = bf(torch.where(big0, w0, w1))
    b = 1.0 - a
    assert torch.equal(bf(b), b)
    return torch.where(big0, a, b), torch.where(big0, b, a)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=24)
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--trans", default="camera")
    ap.add_argument("--hoist3", action="store_true", help="also hoist the 32^3 level")
    a = ap.parse_args()
    torch.manual_seed(0)
    inp = synth.make_inputs(seed=synth.SEED, B=1, N=8, size="full", trans=a.trans)
    res = a.res
    ax = np.linspace(-0.5, 0.5, res)
    rng = np.random.default_rng(1)
    lines = rng.integers(0, res * res, size=a.lines)
    pts = []
    for ln in lines:
        ix, iy = divmod(int(ln), res)
        for iz in range(res):
            pts.append((ax[ix], ax[iy], ax[iz]))
    pts = torch.tensor(np.array(pts), dtype=torch.float32).unsqueeze(0)          # (1, N, 3)
    with torch.no_grad():
        ref = O.list_query(inp.maps, inp.vols, inp.trans_mat, pts, inp.weights)[0]
    N = pts.shape[1]

    # ---- operands of the bf16 path ----
    W0 = inp.weights["fc.fc_0.weight"].squeeze(-1)                               # (512, 3610) reference column order
    b0 = inp.weights["fc.fc_0.bias"]
    vol_ch = [v.shape[1] for v in inp.vols]
    cum = np.concatenate([[0], np.cumsum(vol_ch)])
    nvox = 7 * int(cum[-1])

    def w0_block(level, d):                                                      # columns (cum+c)*7+d, c over the level
        cols = torch.tensor([(int(cum[level]) + c) * 7 + d for c in range(vol_ch[level])])
        return bf(W0[:, cols])

    maps_cl = bf(O.prepare_maps(inp.maps))[0]                                    # (S,S,1024)
    S = maps_cl.shape[0]
    W0m = bf(W0[:, nvox:nvox + 1024])
    P = bf(maps_cl.reshape(-1, 1024) @ W0m.t())                                  # (S*S, 512) projected map, bf16
    vols_cl = [bf(O.to_channels_last_vol(v))[0] for v in inp.vols]               # (R,R,R,C)
    hoisted = [5, 4] + ([3] if a.hoist3 else [])
    PV = {h: [bf(vols_cl[h].reshape(-1, vol_ch[h]) @ w0_block(h, d).t()).reshape(*vols_cl[h].shape[:3], 512) for d in range(7)]
          for h in hoisted}
    disp = O.displacements()

    q = pts[0][:, [2, 1, 0]] * 2                                                 # (N,3): [0]->W (walk axis), [1]->H, [2]->D
    results = {}
    for mode in ("exact", "rn", "unit"):
        add = torch.zeros(N, 512)
        # voxel levels: G rows per (line, class, node), then the z-lerp with bf16 weights
        for h in hoisted:
            R = vols_cl[h].shape[0]
            for cls, dl in ((0, [0, 3, 4, 5, 6]), (1, [1]), (2, [2])):
                G = torch.zeros(N, R, 512)                                       # per point (wasteful but simple): its line's table
                for d in dl:
                    z0, z1, wz0, wz1 = axis(q[:, 2] + disp[d, 2], R)
                    y0, y1, wy0, wy1 = axis(q[:, 1] + disp[d, 1], R)
                    for zi, wz in ((z0, wz0), (z1, wz1)):
                        for yi, wy in ((y0, wy0), (y1, wy1)):
                            G += (wy * wz).view(N, 1, 1) * PV[h][d][zi, yi]      # (N,R,512)
                G = bf(G)
                shift = 0.0 if cls == 0 else (-DISP if cls == 1 else DISP)
                i0, i1, w0, w1 = axis(q[:, 0] + shift, R)
                w0q, w1q = quant_pair(w0, w1, mode)
                idx = torch.arange(N)
                add += w0q.view(N, 1) * G[idx, i0] + w1q.view(N, 1) * G[idx, i1]
        # pixel taps
        xy, _ = O.localise(q.unsqueeze(0), inp.trans_mat)
        xy = xy[0]
        half = (S - 1) / 2.0
        g = (xy - half) / half
        ixp = ((g[:, 0] + 1.0) / 2.0) * (S - 1)
        iyp = ((g[:, 1] + 1.0) / 2.0) * (S - 1)
        nan = torch.isnan(ixp) | torch.isnan(iyp)
        ixp = torch.where(nan, torch.zeros_like(ixp), ixp)
        iyp = torch.where(nan, torch.zeros_like(iyp), iyp)
        x0, y0 = ixp.floor(), iyp.floor()
        wx1, wx0 = ixp - x0, (x0 + 1) - ixp
        wy1, wy0 = iyp - y0, (y0 + 1) - iyp
        wx0q, wx1q = quant_pair(wx0, wx1, mode)
        wy0q, wy1q = quant_pair(wy0, wy1, mode)
        Pm = P.view(S, S, 512)
        for (xi, wx) in ((x0, wx0q), (x0 + 1, wx1q)):
            for (yi, wy) in ((y0, wy0q), (y0 + 1, wy1q)):
                ok = (xi <= S - 1) & (yi <= S - 1) & ~nan
                w = wx * wy
                if mode != "exact":
                    w = bf(w)
                w = w * ok.float()
                add += w.view(N, 1) * Pm[yi.clamp(max=S - 1).long(), xi.clamp(max=S - 1).long()]
        # remaining columns: bf16 features x bf16 weights, fp32 accumulate
        rest_levels = [l for l in range(6) if l not in hoisted]
        acc = add + b0
        for l in rest_levels:
            f = O.gather3d(vols_cl[l].unsqueeze(0), (q.unsqueeze(0) + disp.view(7, 1, 3)).reshape(1, 7 * N, 3)).reshape(7, N, vol_ch[l])
            for d in range(7):
                acc += bf(f[d]) @ w0_block(l, d).t()
        acc += bf(q) @ bf(W0[:, nvox + 1024:]).t()
        h1 = bf(torch.relu(acc))
        w = inp.weights
        h2 = bf(torch.relu(h1 @ bf(w["fc.fc_1.weight"].squeeze(-1)).t() + w["fc.fc_1.bias"]))
        h3 = torch.relu(h2 @ bf(w["fc.fc_2.weight"].squeeze(-1)).t() + w["fc.fc_2.bias"])
        sdf = h3 @ w["fc.fc_out.weight"].squeeze(-1).t() + w["fc.fc_out.bias"]
        err = (sdf[:, 0] - ref).abs()
        results[mode] = (err.max().item(), err.median().item(), err.mean().item())
        print(f"weights {mode:6s}: max|dSDF| {err.max().item():.3e}  median {err.median().item():.3e}  mean {err.mean().item():.3e}"
              f"   (scaled SDF, tolerance 2e-2; /sdf_scale for the bench units)")
    # statistics a kernel plan needs: distinct pixel cells / voxel nodes per 128-step tile
    cells = (y0.long() * S + x0.long()).view(a.lines, res)
    per_tile = []
    for ln in range(a.lines):
        for t0 in range(0, res, 128):
            c = cells[ln, t0:t0 + 128]
            per_tile.append(int((c[1:] != c[:-1]).sum()) + 1)
    print(f"pixel cells per 128-step tile: mean {np.mean(per_tile):.1f} max {np.max(per_tile)}")


if __name__ == "__main__":
    main()
