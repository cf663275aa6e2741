#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout 300 python - <<'PY'
import os, torch
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
print(f"lib={os.environ.get('LIST_B200_LIB','default')[-12:]} skip={os.environ.get('LIST_B200_FUSED_SKIP')} full grid {4*ms:.1f} ms")
PY
}
export LIST_B200_FUSED=1
for skip in 7; do
  LIST_B200_FUSED_SKIP=$skip run
  LIST_B200_FUSED_SKIP=$skip LIST_B200_LIB=$PWD/learning-implicitly-from-spatial-transformers-network_b200/liblist_b200_nb4.so run
done 2>&1 | grep -E "lib=|rror" | tee gpurun_out/fused_probe2.log
