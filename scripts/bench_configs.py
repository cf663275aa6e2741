"""Timings of the other BASELINE.json configurations through the same C ABI (one JSON line each):
  cfg-2  training shape: 8 images x 2048 sampled queries, fp32, forward + backward (list_sdf_fwd / list_sdf_bwd
         behind one torch.autograd.Function) with the reference's SDF loss
  cfg-3  1 image, 128^3 dense grid, 1 GPU (fp32 parity mode and bf16)
  cfg-5  batched serving: 8 images per GPU x 128^3 grids, bf16 (the per-GPU share of 64 images on 8 B200)
bench.py (cfg-4, 256^3) stays the driver's benchmark; this script documents the rest."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from list_b200 import hotpath, synth  # noqa: E402

DEV = torch.device("cuda:0")


def timed(fn, warm=3, steps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def cfg2(return_step=False):
    B, N, scale = 8, 2048, 10.0
    inp = synth.make_inputs(seed=synth.SEED, B=B, N=N, size="full", trans="camera", points="training")
    _, sdf_gt = synth.training_points(B, N, torch.Generator().manual_seed(synth.SEED + 1000))
    g = inp.to(DEV)
    gt = sdf_gt.to(DEV)
    maps = [m.clone().requires_grad_(True) for m in g.maps]
    vols = [v.clone().requires_grad_(True) for v in g.vols]
    T = g.trans_mat.clone().requires_grad_(True)
    w = {k: v.clone().requires_grad_(True) for k, v in g.weights.items()}
    leaves = [*maps, *vols, T, *w.values()]

    def step():
        for t in leaves:
            t.grad = None
        maps_cl = hotpath.prep_maps_autograd(maps)                     # as models.LIST.forward
        vols_cl = [hotpath.prep_volume_autograd(v) for v in vols]
        sdf = hotpath.query_sdf_autograd(g.points, T, maps_cl, vols_cl, w, raw=True)
        loss = ((gt * scale - sdf) ** 2).sum(-1).mean()             # reference losses.py:21-27
        loss.backward()
        return loss

    def hot_only():                                                    # the hot path alone (no upsample / layout glue)
        with torch.no_grad():
            ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, "fp32")
            kw = hotpath.prepare_weights(g.weights, ctx.layout, "fp32")
            return hotpath.query_sdf(ctx, kw, g.points)
    if return_step:
        return step
    ms = timed(step)
    ms_fwd = timed(hot_only)
    loss = float(step().item())
    return {"config": "cfg-2: 8 images x 2048 sampled queries (0.45/0.44/0.1, sdf_scale 10), fwd+bwd, fp32, 1 B200",
            "ms_per_step": ms, "queries_per_sec": B * N / (ms * 1e-3), "loss": loss,
            "ms_fwd_inference_path": ms_fwd,
            "note": "step = differentiable prep kernels (upsample + layouts) + list_sdf_fwd kernels + SDF loss + list_sdf_bwd kernels "
                    "(gradients w.r.t. MLP weights, 6 volumes, 5 maps, trans_mat)"}


def grid_cfg(name, B, res, mode):
    inp = synth.make_inputs(seed=synth.SEED, B=B, N=8, size="full", trans="camera")
    g = inp.to(DEV)
    ctx = hotpath.prepare_context(g.maps, g.vols, g.trans_mat, mode)
    kw = hotpath.prepare_weights(g.weights, ctx.layout, mode)
    total = res ** 3
    chunk = 1048576 if mode == "bf16" else 131072
    out = torch.empty(B, total, device=DEV, dtype=torch.float32)
    cs, wsn = ctx.struct(), kw.struct()
    ws = hotpath._workspace(cs, wsn, chunk, DEV, res)
    ms = timed(lambda: hotpath.grid_sdf(ctx, kw, res, 0, total, 10.0, chunk, out=out, workspace=ws))
    return {"config": name, "dtype": mode, "images": B, "grid_res": res, "ms_per_step": ms,
            "queries_per_sec": B * total / (ms * 1e-3), "checksum": float(out.double().sum().item())}


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg3", "cfg5"]
    if "cfg2" in which:
        print(json.dumps(cfg2()), flush=True)
    if "cfg3" in which:
        print(json.dumps(grid_cfg("cfg-3: 1 image, 128^3 dense SDF grid, 1 B200", 1, 128, "bf16")), flush=True)
        print(json.dumps(grid_cfg("cfg-3: 1 image, 128^3 dense SDF grid, 1 B200 (fp32 parity mode)", 1, 128, "fp32")), flush=True)
    if "cfg5" in which:
        print(json.dumps(grid_cfg("cfg-5: batched serving, 8 images per GPU x 128^3 grids (64 images on 8 B200), bf16", 8, 128, "bf16")), flush=True)
