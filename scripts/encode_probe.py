"""GPU probe: cost of the per-image stage (stock PyTorch encoders + prep kernels), which is outside the per-query metric."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from list_b200.network import models
from oracle.ref_import import RefConfig

dev = "cuda:0"
torch.manual_seed(0)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for B in (1, 8):
    cfg = RefConfig(); cfg.train_batch_size = B
    net = models.LIST(cfg).to(dev).eval()
    img = torch.rand(B, 3, 224, 224, device=dev)
    with torch.no_grad():
        t_all = timed(lambda: net.encode(img, dtype="bf16"))
        t_img = timed(lambda: (net.im_encoder(img), net.im_encoder2(img)))
        occ = torch.zeros(B, 128, 128, 128, device=dev); occ[:, 40:90, 40:90, 40:90] = 1
        t_vox = timed(lambda: net.vox_encoder(occ))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            t_vox16 = timed(lambda: net.vox_encoder(occ))
        net_cl = net.vox_encoder.to(memory_format=torch.channels_last_3d)
        occ_cl = occ
        with torch.autocast("cuda", dtype=torch.bfloat16):
            t_vox16cl = timed(lambda: net_cl(occ_cl))
    print(f"B={B}: encode (all per-image stages + prep) {t_all:.2f} ms | two ResNet-18 {t_img:.2f} ms | vox_encoder fp32/TF32 {t_vox:.2f} ms, "
          f"bf16 autocast {t_vox16:.2f} ms, bf16 + channels_last_3d {t_vox16cl:.2f} ms")
    del net
