"""GPU debug aid: checks the stages of the hoisted-fc_0 path separately (projection GEMMs, addend parts)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from list_b200 import hotpath, synth

DEV = "cuda:0"
size = sys.argv[1] if len(sys.argv) > 1 else "small"
B = 2 if size == "small" else 1
inp = synth.make_inputs(seed=21, B=B, N=8, size=size, trans="camera")
g = inp.to(DEV)
ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "bf16")
kw = hotpath.prepare_weights(g.weights, ctx.layout, "bf16")
ctx32 = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "fp32")
kw32 = hotpath.prepare_weights(g.weights, ctx32.layout, "fp32")
hs = hotpath.HoistedState(ctx, kw)
torch.cuda.synchronize()
up = lambda x: (x + 255) // 256 * 256
S = ctx.maps_cl.shape[1]
off = up(512 * hs.k_h * 2)
n = B * S * S * 512
pmap = hs.buf[off: off + n * 2].view(torch.bfloat16).view(B * S * S, 512).float()
off += up(n * 2)
W = kw.w0.float()
want = ctx.maps_cl.view(-1, 1024).float() @ W[:, :1024].t()
print("pmap   max err %.3e  (max |want| %.3f)" % ((pmap - want).abs().max().item(), want.abs().max().item()))
lay = ctx.layout
for l in (5, 4):
    R, Cc = ctx.vols_cl[l].shape[1], ctx.vols_cl[l].shape[4]
    n = 7 * B * R ** 3 * 512
    pv = hs.buf[off: off + n * 2].view(torch.bfloat16).view(7, B * R ** 3, 512).float()
    off += up(n * 2)
    V = ctx.vols_cl[l].view(-1, Cc).float()
    for d in range(7):
        c0 = lay.vol_off[l] + d * Cc
        want = V @ W[:, c0:c0 + Cc].t()
        print("pvol L%d d%d max err %.3e (max |want| %.3f)" % (l, d, (pv[d] - want).abs().max().item(), want.abs().max().item()))
res, begin, count = (32, 0, 32 ** 3) if size == "small" else (256, 256 * 256 * 100, 256 * 64)
Xh = hs.gather_grid(0, res, begin, count).float()
full32 = hotpath.gather_grid_features(ctx32, 0, res, begin, count)
W32 = kw32.w0
parts = {"2d": (0, 1024), "L5": (1024, 1920), "L4": (1920, 2816)}
comp = {k: full32[:, a:b] @ W32[:, a:b].t() for k, (a, b) in parts.items()}
tot = sum(comp.values())
add = Xh[:, :512]
print("addend vs total: %.3e (max %.3f)" % ((add - tot).abs().max().item(), tot.abs().max().item()))
for k in comp:
    print("addend - (total - %s): %.3e ; addend - %s only: %.3e" % (k, (add - (tot - comp[k])).abs().max().item(), k, (add - comp[k]).abs().max().item()))
err = (add - tot).abs()
rowerr = err.max(dim=1).values
bad = (rowerr > 0.05).nonzero().flatten()
print("bad rows:", bad.numel(), "of", count, "first:", bad[:40].tolist())
