"""Phase timeline of grid_tc_kernel (CTA 0, first 16 tile pairs) and per-kernel times on a range of the 256^3 grid."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from list_b200 import hotpath, synth                 # noqa: E402


def main():
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 4194304
    trans = sys.argv[3] if len(sys.argv) > 3 else "camera"
    dev = torch.device("cuda:0")
    inp = synth.make_inputs(seed=synth.SEED, B=1, N=8, size="full", trans=trans).to(dev)
    ctx = hotpath.prepare_context(inp.maps, inp.vols, inp.trans_mat, "bf16")
    kw = hotpath.prepare_weights(inp.weights, ctx.layout, "bf16")
    ls = hotpath.LineTableState(ctx, kw)
    begin = (res ** 3 // 2 // (res * res)) * res * res
    rows = min(rows, res ** 3 - begin)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    stats = torch.zeros(2, device=dev, dtype=torch.int64)
    G = ls.table(0, res, begin, rows)
    X = ls.rest(0, res, begin, rows)
    plan = ls.plan(0, res, begin, rows, G)
    for rep in range(3):
        stats.zero_()
        ev[0].record()
        ls.table(0, res, begin, rows, out=G)
        ev[1].record()
        ls.rest(0, res, begin, rows, out=X)
        ev[2].record()
        ls.plan(0, res, begin, rows, G, out=plan)
        ev[3].record()
        sdf, tr = ls.evaluate(res, begin, rows, X, plan, 10.0, trace=True, stats=stats)
        ev[4].record()
        torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    pairs, ksteps = (int(x) for x in stats.cpu())
    chunks = ksteps / 4
    print(f"res {res} rows {rows} T={trans}: lines {t[0]:.3f} ms, rest {t[1]:.3f} ms, plan {t[2]:.3f} ms, grid_tc {t[3]:.3f} ms "
          f"({rows / t[3] / 1e3:.1f} M rows/s); {chunks / max(pairs, 1):.2f} interpolation chunks per tile pair")
    tr = tr.cpu().numpy().astype(np.float64)
    names = ["start", "fc0 issued", "fc1 start", "fc1 issued", "fc2 start", "fc2 issued", "fc0 done", "ep0 done", "fc1 done",
             "ep1 done", "fc2 done", "ep2 done", "-", "I issued", "plan loaded", "I filled",
             "loop top", "wall ns", "plan0 done", "grant c0", "A done", "cp issued", "cp landed", "arrived"]
    t0 = tr[:, 0:1]
    rel = tr - t0
    np.set_printoptions(linewidth=200, suppress=True)
    print("cycles relative to the tile's start (rows = recorded tiles 2..15 of CTA 0):")
    print("  " + "  ".join(f"{n:>10s}" for n in names))
    for i in range(2, 16):
        print("  " + "  ".join(f"{rel[i, j]:10.0f}" for j in range(len(names))))
    stride = int(os.environ.get("LIST_B200_TRACE_STRIDE", "1"))
    dcyc, dns = tr[15, 0] - tr[2, 0], tr[15, 17] - tr[2, 17]
    print(f"SM clock under load (clock64 / globaltimer between recorded tiles 2 and 15): {dcyc / max(dns, 1):.3f} GHz")
    print("tile period (cycles):", np.diff(tr[2:16, 0]) / stride)


if __name__ == "__main__":
    main()
