"""GPU tests of the host-side mirror of the reference's module / model / executor API: the two hot
modules called with the reference's own signatures, `LIST.forward` in inference and training mode, and the
executor's dense-grid evaluation + mesh extraction, each against the ATen-op oracle fed with the SAME per-image
tensors (so that the comparison isolates the hot path from the encoders' TF32 convolutions)."""
import numpy as np
import pytest
import torch

from list_b200 import hotpath, synth
from list_b200.network import executors, models, modules
from oracle import mcubes_oracle as MC
from oracle import ref_port as P
from oracle.ref_import import RefConfig

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Cfg(RefConfig):
    sdf_scale = 10.0
    grid_res = 32
    device = DEV


@pytest.fixture(scope="module")
def net():
    torch.manual_seed(333)
    return models.LIST(Cfg()).to(DEV).eval()


def sd_weights(net):
    return {k[len("sdf_decoder."):]: v.detach().cpu() for k, v in net.state_dict().items() if k.startswith("sdf_decoder.fc.")}


def test_hot_modules_with_the_reference_signatures():
    """PerceptualPooling.forward / VoxelDecoder2.forward exactly as executors.py:219-223 calls them."""
    inp = synth.make_inputs(seed=31, B=2, N=300, size="small", trans="camera")
    g = inp.to(DEV)
    q = (g.points[:, :, [2, 1, 0]] * 2).contiguous()                      # executors.py:217-218
    pool = modules.PerceptualPooling()
    dec = modules.VoxelDecoder2(inp.feature_size, 256).to(DEV)
    dec.load_state_dict({k: v for k, v in g.weights.items()})
    with torch.no_grad():                                                  # executors.py:199 (the test loop runs under no_grad)
        percep = pool(g.maps, q, g.trans_mat)
        assert pool(g.maps, q[:, :100].contiguous(), g.trans_mat).shape == (2, 1024, 1, 100)     # second chunk: cached context
    assert percep.shape == (2, 1024, 1, 300)
    qc = q.cpu()
    ref_p = P.perceptual_pooling(inp.maps, qc, inp.trans_mat)
    assert (percep.cpu() - ref_p).abs().max().item() <= 2e-5
    # the stand-alone modules run the kernels on detached tensors: under autograd they refuse instead of training nothing
    with pytest.raises(RuntimeError, match="inference-only"):
        dec(q, g.vols, percep.reshape(2, -1, 300))
    with pytest.raises(RuntimeError, match="inference-only"):
        pool([m.clone().requires_grad_(True) for m in g.maps], q, g.trans_mat)
    with torch.no_grad():
        sdf = dec(q, g.vols, percep.reshape(2, -1, 300))
    assert sdf.shape == (2, 300)
    ref = P.voxel_decoder2(qc, inp.vols, ref_p.reshape(2, -1, 300), inp.weights)
    assert (sdf.cpu() - ref).abs().max().item() <= 1e-4


def test_list_forward_inference_and_query_api(net):
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 3, 224, 224, generator=g).to(DEV)
    pts = (torch.rand(1, 1000, 3, generator=g) - 0.5).to(DEV)
    with torch.no_grad():
        maps, vols, T = net.per_image(img)
        occ, sdf = net(img, pts)
        ctx = net.encode(img)
        sdf2 = net.query(ctx, pts)
        ctx16 = net.encode(img, dtype="bf16")
        sdf16 = net.query(ctx16, pts)
        ref = P.list_query([m.cpu() for m in maps], [v.cpu() for v in vols], T.cpu(), pts.cpu(), sd_weights(net))
    assert occ.shape == (1, 1, 128, 128, 128) and sdf.shape == (1, 1000)
    assert torch.equal(sdf, sdf2)
    assert (sdf.cpu() - ref).abs().max().item() <= 1e-4
    assert (sdf16.cpu() - ref).abs().max().item() <= 2e-2


def test_executor_grid_and_mesh(net):
    ex = executors.LIST(Cfg(), net)
    g = torch.Generator().manual_seed(6)
    batch = {"rgb_image": torch.rand(1, 3, 224, 224, generator=g)}
    vals, ctx = ex.predict_grid(batch)
    assert vals.shape == (32, 32, 32) and vals.dtype == np.float32
    with torch.no_grad():
        maps, vols, T = net.per_image(batch["rgb_image"].to(DEV), unsqueeze_dim=0)
    ref = P.dense_grid_sdf([m.cpu() for m in maps], [v.cpu() for v in vols], T.cpu(), sd_weights(net), 32, 10.0, chunk=8192)
    assert np.abs(vals - ref).max() <= 1e-4
    (mesh, occ, occ_pred), scores = ex.test(batch)
    v_ref, t_ref = MC.generate_mesh(vals, -0.5, 0.5)
    assert scores == {} and occ_pred.shape[-3:] == (128, 128, 128)
    assert np.asarray(mesh.vertices).shape == v_ref.shape and np.array_equal(np.asarray(mesh.faces), t_ref)
    if len(v_ref):
        assert np.abs(np.asarray(mesh.vertices) - v_ref).max() <= 1e-5


def test_list_forward_training_mode_reaches_every_trainable_stage():
    torch.manual_seed(7)
    cfg = Cfg()
    cfg.train_batch_size = 2
    net = models.LIST(cfg).to(DEV).train()
    g = torch.Generator().manual_seed(8)
    img = torch.rand(2, 3, 224, 224, generator=g).to(DEV)
    pts, gt = synth.training_points(2, 256, g)
    ex = executors.LIST(cfg, net)
    batch = {"rgb_image": img, "points": pts, "values": gt, "occ": (torch.rand(2, 1, 128, 128, 128, generator=g) > 0.99).float()}
    pred, loss = ex.train(batch, calc_loss=True)
    total = sum(v for k, v in loss.items() if "ignore" not in k)
    total.backward()
    assert pred[1].shape == (2, 256) and torch.isfinite(total)
    for name in ("sdf_decoder.fc.fc_0.weight", "sdf_decoder.fc.fc_out.bias", "vox_encoder", "im_encoder2", "spatial_transformer"):
        named = [(n, p.grad) for n, p in net.named_parameters() if n.startswith(name) and p.requires_grad]
        missing = [n for n, gr in named if gr is None]
        bad = [n for n, gr in named if gr is not None and not torch.isfinite(gr).all()]
        assert named and not bad, (name, bad)
        # parameters of layers the reference constructs but never calls get no gradient there either
        assert len(missing) < len(named), (name, missing)
        assert any(gr is not None and gr.abs().sum() > 0 for _, gr in named), name
        if missing:
            print(f"{name}: no gradient for {missing}")


def test_kernel_weight_cache_under_threaded_replicas():
    """nn.DataParallel replicas share the module's __dict__ shallowly, so every replica thread sees the same cache dict
    (reference train.py:126, test.py:62 wrap the model that way).  Threads asking for different (device, dtype) slots --
    here the two compute dtypes on one device -- must each get the object built for THEIR slot from THEIR parameters."""
    import copy
    import threading
    from list_b200.network import modules as M
    torch.manual_seed(0)
    dec = M.VoxelDecoder2(3610, 256).cuda()
    layout = hotpath.feature_layout(1024, [1, 16, 32, 64, 128, 128])
    replicas = [copy.copy(dec) for _ in range(4)]                  # shallow: the cache dict is shared, like DataParallel's replicate
    assert all(r._cache is dec._cache for r in replicas)
    errors = []

    def worker(rep, dtype, rounds=25):
        try:
            want = torch.bfloat16 if dtype == "bf16" else torch.float32
            ref_w1 = rep.fc["fc_1"].weight.detach().squeeze(-1).to(want)
            for _ in range(rounds):
                kw = rep.kernel_weights(layout, dtype)
                if kw.w0.dtype != want or not kw.w0.is_cuda:
                    errors.append((dtype, str(kw.w0.dtype)))
                elif not torch.equal(kw.w1.reshape(ref_w1.shape), ref_w1):
                    errors.append((dtype, "w1 mismatch"))
        except Exception as e:                                     # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(replicas[i], "bf16" if i % 2 else "fp32")) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    assert not errors, errors[:3]
