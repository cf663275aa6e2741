"""GPU probe: how much do the MLP and the gather kernels slow each other down when they share the SMs?
Runs each kernel alone, then the MLP on one stream concurrently with a gather kernel on another."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from list_b200 import hotpath, synth

dev = torch.device("cuda:0")
inp = synth.make_inputs(seed=synth.SEED, B=1, N=8, size="full", trans="camera")
g = inp.to(dev)
ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "bf16")
kw = hotpath.prepare_weights(g.weights, ctx.layout, "bf16")
hs = hotpath.HoistedState(ctx, kw)
res, rows = 256, int(os.environ.get("ROWS", 1048576))
XA = hs.gather_grid(0, res, 0, rows)                 # rows the MLP reads
XB = torch.empty_like(XA)                            # rows the gather writes
torch.cuda.synchronize()
sa, sb = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)


def ev():
    return torch.cuda.Event(enable_timing=True)


def alone(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


mlp = lambda: hs.mlp(XA, 10.0)
add = lambda: hs.gather_grid(0, res, rows, rows, parts=1, out=XB)
rest = lambda: hs.gather_grid(0, res, rows, rows, parts=2, out=XB)
t_mlp, t_add, t_rest = alone(mlp), alone(add), alone(rest)
print(f"alone: mlp {t_mlp:.3f} ms, addend {t_add:.3f} ms, rest {t_rest:.3f} ms")


def corun(other, name, n=5):
    res_ = []
    for _ in range(n + 1):
        torch.cuda.synchronize()
        a0, a1, b0, b1 = ev(), ev(), ev(), ev()
        with torch.cuda.stream(sa):
            a0.record(); mlp(); a1.record()
        with torch.cuda.stream(sb):
            b0.record(); other(); b1.record()
        torch.cuda.synchronize()
        res_.append((a0.elapsed_time(a1), b0.elapsed_time(b1), max(a0.elapsed_time(a1), a0.elapsed_time(b1))))
    r = res_[1:]
    m = [sum(x[i] for x in r) / len(r) for i in range(3)]
    print(f"mlp || {name}: mlp {m[0]:.3f} ms, {name} {m[1]:.3f} ms, both done after {m[2]:.3f} ms "
          f"(serial would be {t_mlp + (t_add if name == 'addend' else t_rest):.3f})")


corun(add, "addend")
corun(rest, "rest")
