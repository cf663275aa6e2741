"""csrc/tgemm.cu: the fp32-accurate tensor-core GEMM (three TF32 products per fp32 product) of the fp32 path, against
fp64 torch.  Plain TF32 would be ~1e-3 relative here; the bound below is what keeps the 1e-4 SDF parity gate."""
import os

import pytest
import torch

from list_b200 import hotpath, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 512, 3648), (512, 3648, 4096), (1000, 256, 512), (77, 512, 36)])
def test_gemm_matches_fp64(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    b = torch.randn(N, K, generator=g).to(DEV)
    ref = a.double() @ b.double().t()
    got = hotpath.gemm_f32_tc(a, b)
    scale = ref.abs().max().item()
    err = (got.double() - ref).abs().max().item()
    tf32 = 2.0 ** -11 * scale
    print(f"M{M} N{N} K{K}: max err {err:.3e} = {err / scale:.2e} of max |C| (plain TF32 ~ {tf32:.1e})")
    assert err <= 6e-5 * scale
    acc = hotpath.gemm_f32_tc(a, b, out=torch.ones(M, N, device=DEV), accumulate=True)
    assert ((acc - 1).double() - ref).abs().max().item() <= 6e-5 * scale


def test_fp32_mlp_on_tensor_cores_equals_the_ffma_path(monkeypatch):
    """Inference forward of the fp32 mode: 3xTF32 GEMMs (default) against the FFMA GEMMs (LIST_B200_F32_TC=0)."""
    inp = synth.make_inputs(seed=3, B=2, N=1500, size="small").to(DEV)
    ctx = hotpath.prepare_context(inp.maps, inp.vols, inp.trans_mat, "fp32")
    kw = hotpath.prepare_weights(inp.weights, ctx.layout, "fp32")
    X = hotpath.gather_features(ctx, inp.points)
    tc = hotpath.mlp(kw, X)
    monkeypatch.setenv("LIST_B200_F32_TC", "0")
    ffma = hotpath.mlp(kw, X)
    train = hotpath.mlp(kw, X, train=True)
    assert torch.equal(ffma, train)                     # the training forward always accumulates on the FFMA pipe
    assert (tc - ffma).abs().max().item() <= 2e-5
