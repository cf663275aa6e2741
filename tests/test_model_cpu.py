"""CPU tests of the host-side mirror of the reference's module API: same attribute names and
state_dict keys as `network.models.LIST`, the per-image stage (stock PyTorch) equal to the
reference's when both hold the same weights, device-side voxelisation equal to the cKDTree one."""
import numpy as np
import pytest
import torch

from list_b200.network import executors, models, modules
from oracle import ref_import
from oracle.ref_import import RefConfig

HOT_PATH_ATTRS = ["im_encoder", "im_encoder2", "point_decoder", "point_mlp_coarse", "spatial_transformer",
                  "create_occ", "vox_encoder", "percep_pooling", "sdf_decoder"]       # executors.py:200-223


@pytest.fixture(scope="module")
def net():
    torch.manual_seed(333)
    return models.LIST(RefConfig()).eval()


def test_list_has_the_reference_attributes_and_checkpoint_keys(net):
    for name in HOT_PATH_ATTRS:
        assert hasattr(net, name), name
    sd = net.state_dict()
    shapes = {"sdf_decoder.fc.fc_0.weight": (512, 3610, 1), "sdf_decoder.fc.fc_1.weight": (256, 512, 1),
              "sdf_decoder.fc.fc_2.weight": (256, 256, 1), "sdf_decoder.fc.fc_out.weight": (1, 256, 1),
              "sdf_decoder.fc.fc_0.bias": (512,), "sdf_decoder.fc.fc_out.bias": (1,)}      # modules.py:196-200
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == s
    assert sum(p.numel() for p in net.sdf_decoder.parameters()) == 2_046_209             # 2.05 M
    assert sum(p.numel() for p in net.parameters()) == 105_236_071


def test_plugin_lookup_by_dotted_path():
    """The reference selects model and executor with utils.get_class (utils.py:20-26, test.py:60,95)."""
    def get_class(kls):
        parts = kls.split(".")
        m = __import__(".".join(parts[:-1]))
        for comp in parts[1:]:
            m = getattr(m, comp)
        return m
    name = "list_b200.network.models.LIST"
    assert get_class(name) is models.LIST
    assert get_class(name.replace("model", "executor")) is executors.LIST


def test_hot_path_modules_refuse_cpu_tensors(net):
    pool = modules.PerceptualPooling()
    maps = [torch.zeros(1, 8, 4, 4)]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pool(maps, torch.zeros(1, 5, 3), torch.zeros(1, 4, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.rand(1, 3, 224, 224), torch.zeros(1, 16, 3))


def test_create_occ_equals_kdtree_voxelisation(net):
    from scipy.spatial import cKDTree
    from oracle.list_oracle import create_grid_points_from_bounds
    R = 16
    net_small = models.LIST.__new__(models.LIST)
    torch.nn.Module.__init__(net_small)
    net_small.vox_res, net_small.bb_min, net_small.bb_max = R, -0.5, 0.5
    pc = (torch.rand(2, 500, 3, generator=torch.Generator().manual_seed(0)) - 0.5) * 1.2    # some outside the box
    occ = models.LIST.create_occ(net_small, pc)
    tree = cKDTree(create_grid_points_from_bounds(-0.5, 0.5, R))
    ref = torch.zeros(2, R ** 3)
    for b in range(2):                                     # reference models.py:107-109
        _, idx = tree.query(pc[b].numpy())
        ref[b][idx] = 1
    assert torch.equal(occ, ref.view(2, R, R, R))


@pytest.mark.skipif(not ref_import.available(), reason="needs /root/reference (build container only)")
def test_per_image_stage_equals_reference_with_shared_weights(net):
    _, ref_models = ref_import.load()
    torch.manual_seed(0)
    ref = ref_models.LIST(RefConfig()).eval()
    missing, unexpected = net.load_state_dict(ref.state_dict(), strict=True), None       # identical key sets
    img = torch.rand(1, 3, 224, 224, generator=torch.Generator().manual_seed(333))
    with torch.no_grad():
        maps, vols, T = net.per_image(img)
        feat_g, _ = ref.im_encoder(img)
        feat_g2, feat_l2 = ref.im_encoder2(img)
        pc = ref.point_decoder([feat_g.unsqueeze(1)])
        fc = torch.max(ref.point_mlp_coarse(pc), -1)[0].reshape(1, -1)
        T_ref = ref.spatial_transformer(torch.cat([fc, feat_g2.reshape(1, -1)], dim=1)).reshape(-1, 4, 3)
        vols_ref = ref.vox_encoder(ref.create_occ(pc))
    assert [tuple(m.shape) for m in maps] == [(1, 64, 224, 224), (1, 64, 112, 112), (1, 128, 56, 56),
                                              (1, 256, 28, 28), (1, 512, 14, 14)]
    assert [tuple(v.shape) for v in vols] == [(1, 1, 128, 128, 128), (1, 16, 128, 128, 128), (1, 32, 64, 64, 64),
                                              (1, 64, 32, 32, 32), (1, 128, 16, 16, 16), (1, 128, 8, 8, 8)]
    for a, b in zip(maps, feat_l2):
        assert torch.allclose(a, b, atol=1e-5)
    assert torch.allclose(T, T_ref, atol=1e-5)
    for a, b in zip(vols, vols_ref):
        assert torch.allclose(a, b, atol=1e-5)
