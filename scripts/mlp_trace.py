"""GPU diagnostic: phase timeline of the hoisted MLP kernel (CTA 0, first 16 tiles), in microseconds."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from list_b200 import _C, hotpath, synth

dev = torch.device("cuda:0")
inp = synth.make_inputs(seed=synth.SEED, B=1, N=8, size="full", trans="camera")
g = inp.to(dev)
ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "bf16")
kw = hotpath.prepare_weights(g.weights, ctx.layout, "bf16")
hs = hotpath.HoistedState(ctx, kw)
rows = 1048576
X = hs.gather_grid(0, 256, 0, rows)
sdf = torch.empty(rows, device=dev)
trace = torch.zeros(16, 12, device=dev, dtype=torch.int64)
ws = kw.struct()
for _ in range(2):
    _C.check(_C.lib().list_mlp_hoisted_trace(C.byref(ws), hs.hoist_cols, X.data_ptr(), X.stride(0), rows, sdf.data_ptr(), 10.0,
                                             trace.data_ptr(), torch.cuda.current_stream().cuda_stream), "trace")
torch.cuda.synchronize()
t = trace.cpu().double()
mhz = float(os.popen("nvidia-smi --query-gpu=clocks.sm --format=csv,noheader,nounits -i 0").read().split()[0])
print("sm clock now", mhz, "MHz (stamps are SM cycles; us below assume 1900 MHz)")
us = (t - t[0, 0]) / 1900.0
names = ["start", "fc0_iss", "fc1_start", "fc1_iss", "fc2_start", "fc2_iss", "fc0_done", "ep0_done", "fc1_done", "ep1_done", "fc2_done", "ep2_done"]
for tile in range(3, 9):
    r = us[tile]
    print(f"tile {tile}: " + " ".join(f"{n}={r[i] - r[0]:6.2f}" for i, n in enumerate(names)) + f" | next start {us[tile + 1][0] - r[0]:6.2f}")
d = us[4:15]
seg = {"fc_0 (start->done)": d[:, 6] - d[:, 0], "ep0": d[:, 7] - d[:, 6], "ep0->fc1 start": d[:, 2] - d[:, 7], "fc_1 (start->done)": d[:, 8] - d[:, 2],
       "ep1": d[:, 9] - d[:, 8], "ep1->fc2 start": d[:, 4] - d[:, 9], "fc_2 (start->done)": d[:, 10] - d[:, 4], "ep2": d[:, 11] - d[:, 10]}
tot = 0
for k, v in seg.items():
    print(f"{k:22s} {v.mean():6.2f} us"); tot += v.mean()
print("sum", round(float(tot), 2), "us; tile period", round(float((us[14, 0] - us[4, 0]) / 10), 2), "us")
