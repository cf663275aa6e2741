"""Small fused-kernel invocation for profiling / debugging (run from the repo root)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from list_b200 import hotpath, synth
os.environ.setdefault("LIST_B200_FUSED", "1")
dev = torch.device("cuda:0")
inp = synth.make_inputs(seed=333, B=1, N=8, size="full", trans="camera").to(dev)
ctx = hotpath.prepare_context(inp.maps, inp.vols, inp.trans_mat, "bf16")
kw = hotpath.prepare_weights(inp.weights, ctx.layout, "bf16")
res = 256; count = int(os.environ.get("COUNT", 256 * 256 * 16))
out = torch.empty(1, count, device=dev)
for _ in range(3): hotpath.grid_sdf(ctx, kw, res, 0, count, 10.0, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); hotpath.grid_sdf(ctx, kw, res, 0, count, 10.0, out=out); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"ok sum={out.float().abs().sum().item():.4f} {ms:.3f} ms for {count} pts -> {ms * 256**3 / count:.1f} ms per 256^3")
