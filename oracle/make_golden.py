"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_import.py) on seeded synthetic
inputs from list_b200.synth.  Run in the build container:

    python -m oracle.make_golden

The fixtures store only (a) the recipe needed to regenerate the inputs from the
seed and (b) the reference's outputs, so they stay small; tests regenerate the
inputs on whatever box they run on (the GPU box has no /root/reference).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from list_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> synth.make_inputs kwargs
CASES = {
    "small_camera_b2":  dict(seed=333, B=2, N=257, size="small", trans="camera"),
    "small_random_b1":  dict(seed=334, B=1, N=513, size="small", trans="random"),
    "full_camera_b1":   dict(seed=333, B=1, N=2048, size="full", trans="camera"),
    "full_random_b1":   dict(seed=335, B=1, N=1024, size="full", trans="random"),
    "train_b2":         dict(seed=336, B=2, N=2028, size="small", trans="camera", points="training"),
}
GRAD_CASES = {
    "grad_small_b2": dict(seed=337, B=2, N=384, size="small", trans="camera", points="training"),
}


def ref_forward(ref_modules, inp, with_features=False):
    pool = ref_modules.PerceptualPooling()
    dec = ref_modules.VoxelDecoder2(inp.feature_size, synth.H_DIM)
    dec.load_state_dict(inp.weights)
    B, N, _ = inp.points.shape
    q = inp.points[:, :, [2, 1, 0]] * 2                      # models.py:91-92
    percep = pool(inp.maps, q, inp.trans_mat).reshape(B, -1, N)
    sdf = dec(q, inp.vols, percep)
    return (sdf, percep, dec) if with_features else sdf


def main():
    torch.set_num_threads(os.cpu_count())
    ref_modules, _ = ref_import.load()
    os.makedirs(GOLDEN, exist_ok=True)
    for name, kw in CASES.items():
        inp = synth.make_inputs(**kw)
        with torch.no_grad():
            sdf, percep, _ = ref_forward(ref_modules, inp, with_features=True)
        np.savez_compressed(
            os.path.join(GOLDEN, f"{name}.npz"),
            recipe=json.dumps(kw), sdf=sdf.numpy(),
            percep_head=percep[:, :, :16].contiguous().numpy(),     # (B,1024,16) spot check
            points_sum=np.float64(inp.points.double().sum().item()),
            weights_sum=np.float64(sum(v.double().sum().item() for v in inp.weights.values())),
        )
        print(name, tuple(sdf.shape), float(sdf.abs().max()))

    for name, kw in GRAD_CASES.items():
        inp = synth.make_inputs(**kw)
        _, sdf_gt = synth.training_points(kw["B"], kw["N"], torch.Generator().manual_seed(kw["seed"] + 1000))
        maps = [m.clone().requires_grad_(True) for m in inp.maps]
        vols = [v.clone().requires_grad_(True) for v in inp.vols]
        T = inp.trans_mat.clone().requires_grad_(True)
        pool = ref_modules.PerceptualPooling()
        dec = ref_modules.VoxelDecoder2(inp.feature_size, synth.H_DIM)
        dec.load_state_dict(inp.weights)
        B, N, _ = inp.points.shape
        q = inp.points[:, :, [2, 1, 0]] * 2
        percep = pool(maps, q, T).reshape(B, -1, N)
        sdf = dec(q, vols, percep)
        sdf_scale = 10.0
        loss = ((sdf_gt * sdf_scale - sdf) ** 2).sum(-1).mean()      # losses.py:21-22
        loss.backward()
        out = dict(recipe=json.dumps(kw), sdf=sdf.detach().numpy(), loss=np.float64(loss.item()),
                   dT=T.grad.numpy(), sdf_scale=np.float64(sdf_scale))
        for k, p in dec.named_parameters():
            g = p.grad
            out["dW_" + k.replace(".", "_")] = g.numpy() if g.numel() <= 70000 else g.flatten()[::97].numpy()
        for i, m in enumerate(maps):
            out[f"dmap{i}_sub"] = m.grad.flatten()[::13].numpy()
            out[f"dmap{i}_sum"] = np.float64(m.grad.double().sum().item())
        for i, v in enumerate(vols):
            out[f"dvol{i}_sub"] = v.grad.flatten()[::7].numpy()
            out[f"dvol{i}_sum"] = np.float64(v.grad.double().sum().item())
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **out)
        print(name, "loss", loss.item())

    # a-8: grid ordering of utils.create_grid_points_from_bounds
    sys.path.insert(0, ref_import.REFERENCE_ROOT)
    import utils as ref_utils
    g5 = ref_utils.create_grid_points_from_bounds(-0.5, 0.5, 5)
    g64 = ref_utils.create_grid_points_from_bounds(-0.5, 0.5, 64)
    g256 = np.linspace(-0.5, 0.5, 256)
    np.savez_compressed(os.path.join(GOLDEN, "grid_points.npz"), g5=g5,
                        g64_f32_head=g64[:130].astype(np.float32), g64_f32_tail=g64[-130:].astype(np.float32),
                        g64_sum=np.float64(g64.astype(np.float32).astype(np.float64).sum(0)),
                        ax64_f32=np.linspace(-0.5, 0.5, 64).astype(np.float32),
                        ax128_f32=np.linspace(-0.5, 0.5, 128).astype(np.float32),
                        ax256_f32=g256.astype(np.float32))
    print("grid ok")


if __name__ == "__main__":
    main()
