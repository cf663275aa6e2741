"""Stand-alone check of the tcgen05 MLP kernel against a bf16-emulating torch reference.
Run in its own process (tests/test_gpu_mlp_tc.py does) so that a device-side trap cannot poison
the pytest process:   python -m tests.tc_check --variant 2 --rows 1000
Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def reference(X, kw):
    """bf16 operands, fp32 accumulation, hidden activations rounded to bf16 (what the kernel does).
    Returns (sdf, h1, h2, h3) with h* the fp32 pre-rounding activations."""
    import torch
    f = lambda t: t.float()
    h1 = torch.relu(f(X) @ f(kw.w0).t() + kw.b0)
    h2 = torch.relu(f(h1.bfloat16()) @ f(kw.w1).t() + kw.b1)
    h3 = torch.relu(f(h2.bfloat16()) @ f(kw.w2).t() + kw.b2)
    return h3 @ kw.w3 + kw.b3, h1, h2, h3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", type=int, default=2)
    ap.add_argument("--rows", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out-div", type=float, default=1.0)
    ap.add_argument("--layers", action="store_true", help="also check the hidden activations layer by layer")
    a = ap.parse_args()
    os.environ["LIST_B200_MLP_VARIANT"] = str(a.variant)
    import torch
    from list_b200 import hotpath, synth
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(a.seed)
    lay = hotpath.feature_layout(1024, [1, 16, 32, 64, 128, 128])
    w = {k: v.to(dev) for k, v in synth.mlp_weights(lay.k_out, g).items()}
    # O(1) outputs: scale the last layers up so that layout bugs cannot hide in small numbers
    w["fc.fc_out.weight"] = w["fc.fc_out.weight"] * 8
    kw = hotpath.prepare_weights(w, lay, "bf16")
    X = torch.zeros(a.rows, lay.k_pad)
    X[:, :lay.k_out] = torch.randn(a.rows, lay.k_out, generator=g)
    X = X.to(dev).bfloat16()
    out = hotpath.mlp(kw, X, a.out_div)
    torch.cuda.synchronize()
    ref, r1, r2, r3 = reference(X, kw)
    ref = ref / a.out_div
    err = (out - ref).abs().max().item()
    res = {"variant": a.variant, "rows": a.rows, "max_err": err, "ref_absmax": ref.abs().max().item(),
           "finite": bool(torch.isfinite(out).all().item())}
    if a.layers:
        out2, h1, h2, h3 = hotpath.mlp_debug(kw, X, a.out_div)
        torch.cuda.synchronize()
        res.update({"h1_err": (h1 - r1).abs().max().item(), "h2_err": (h2 - r2).abs().max().item(),
                    "h3_err": (h3 - r3).abs().max().item(), "h1_absmax": r1.abs().max().item(),
                    "debug_equals_plain": bool(torch.equal(out, out2))})
    print(json.dumps(res))


if __name__ == "__main__":
    main()
