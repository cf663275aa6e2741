import sys, torch
sys.path.insert(0, "/root/repo")
from list_b200 import hotpath, synth
dev = torch.device("cuda:0")
inp = synth.make_inputs(seed=synth.SEED, B=1, N=8, size="full", trans="camera").to(dev)
ctx = hotpath.prepare_context(inp.maps, inp.vols, inp.trans_mat, "bf16")
kw = hotpath.prepare_weights(inp.weights, ctx.layout, "bf16")
res = 256
for shard in (8, 4, 2):
    count = res ** 3 // shard
    begin = (shard // 2) * count
    out = torch.empty(1, count, device=dev)
    for chunk in (262144, 524288, 1048576, 2097152, 4194304):
        if chunk > count: continue
        ws = hotpath._workspace(ctx.struct(), kw.struct(), chunk, dev, res)
        for _ in range(3): hotpath.grid_sdf(ctx, kw, res, begin, count, 10.0, chunk, out=out, workspace=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): hotpath.grid_sdf(ctx, kw, res, begin, count, 10.0, chunk, out=out, workspace=ws)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"1/{shard} of 256^3 ({count} rows), chunk {chunk}: {ms:.3f} ms -> {count / ms / 1e3:.1f} M q/s per GPU")
