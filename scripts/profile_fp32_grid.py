import sys, torch
sys.path.insert(0, "/root/repo")
from torch.profiler import ProfilerActivity, profile
from list_b200 import hotpath, synth
dev = torch.device("cuda:0")
inp = synth.make_inputs(seed=synth.SEED, B=1, N=8, size="full", trans="camera").to(dev)
ctx = hotpath.prepare_context(inp.maps, inp.vols, inp.trans_mat, "fp32")
kw = hotpath.prepare_weights(inp.weights, ctx.layout, "fp32")
res, chunk = 128, 131072
out = torch.empty(1, res ** 3, device=dev)
ws = hotpath._workspace(ctx.struct(), kw.struct(), chunk, dev, res)
f = lambda: hotpath.grid_sdf(ctx, kw, res, 0, res ** 3, 10.0, chunk, out=out, workspace=ws)
for _ in range(2): f()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    f(); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
print("total", sum(e.device_time_total for e in rows) / 1e3, "ms")
for e in rows[:10]: print(f"{e.device_time_total / 1e3:8.3f} ms x{e.count:4d} {e.key[:100]}")
