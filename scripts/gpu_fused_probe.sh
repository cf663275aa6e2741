#!/bin/bash
mkdir -p gpurun_out
for skip in 0 4 1 2 7; do
  echo "== skip=$skip"
  LIST_B200_FUSED_SKIP=$skip timeout 300 python - <<'PY'
import os, torch, time
from list_b200 import hotpath, synth
dev = torch.device("cuda:0")
inp = synth.make_inputs(seed=333, B=1, N=8, size="full", trans="camera").to(dev)
ctx = hotpath.prepare_context(inp.maps, inp.vols, inp.trans_mat, "bf16")
kw = hotpath.prepare_weights(inp.weights, ctx.layout, "bf16")
res = 256; count = res**3 // 4
out = torch.empty(1, count, device=dev)
for _ in range(2): hotpath.grid_sdf(ctx, kw, res, 0, count, 10.0, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2): hotpath.grid_sdf(ctx, kw, res, 0, count, 10.0, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
print(f"skip={os.environ.get('LIST_B200_FUSED_SKIP')} quarter-grid {ms:.2f} ms -> full grid {4*ms:.1f} ms")
PY
done 2>&1 | grep -E "skip=|rror" | tee gpurun_out/fused_probe.log
