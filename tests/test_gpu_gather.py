"""-m gpu: descriptor gather / scatter-add kernels vs the oracle (bit-exact forward, fp32-tolerance backward)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import dev
from read_b200 import ops, _lib as L
from read_b200.texture import PointTexture

pytestmark = pytest.mark.gpu


def test_point_major_roundtrip():
    t = torch.rand(1, 8, 1001, device=dev())
    nd = ops.texture_to_point_major(t)
    assert torch.equal(nd, t[0].t().contiguous())
    assert torch.equal(ops.texture_to_channel_major(nd), t)
    t5 = torch.rand(1, 5, 77, device=dev())
    assert torch.equal(ops.texture_to_point_major(t5), t5[0].t().contiguous())


@pytest.mark.parametrize("name", ["net_64x64_b1", "net_80x48_b2"])
def test_gather_matches_reference_fixture(name):
    g = load_golden(name)
    tex = torch.from_numpy(g["texture"]).to(dev())
    nd = ops.texture_to_point_major(tex)
    ids = torch.from_numpy(g["index0"][:, 0]).to(dev()).contiguous()
    out = ops.gather_from_index(nd, ids, L.FEAT_NCHW_F32)
    np.testing.assert_array_equal(out.cpu().numpy(), g["feat0"])
    nhwc = ops.gather_from_index(nd, ids, L.FEAT_NHWC_F32)
    assert torch.equal(nhwc.permute(0, 3, 1, 2), out)
    bf = ops.gather_from_index(nd, ids, L.FEAT_NHWC_BF16)
    assert torch.equal(bf, nhwc.to(torch.bfloat16))


def test_gather_from_zbuf_equals_gather_from_index():
    from gpu_util import scene_and_cams
    xyz, M = scene_and_cams(30_000, 64, 48, [0, 2])
    d = dev()
    pyr = ops.Pyramid(2, 64, 48, 4, d)
    pyr.clear()
    ops.raster_project(pyr, torch.from_numpy(xyz).to(d), torch.from_numpy(M).to(d))
    nd = torch.rand(30_000, 8, device=d)
    for l in range(4):
        idx, _ = ops.zbuf_resolve(pyr, l)
        a = ops.gather_from_index(nd, idx, L.FEAT_NHWC_F32)
        b = ops.gather_from_zbuf(nd, pyr, l, L.FEAT_NHWC_F32)
        assert torch.equal(a, b)
        assert torch.equal(b[idx == 0], nd[0].expand_as(b[idx == 0]))      # empty pixels read point 0


@pytest.mark.parametrize("act", ["none", "sigmoid", "tanh"])
def test_point_texture_module_forward_backward_vs_oracle(act):
    from oracle import unet_ref
    g = torch.Generator().manual_seed(5)
    N = 4000
    ids = torch.randint(0, N, (2, 1, 40, 56), generator=g).float()
    ids[:, :, :20] = 0                                    # half the pixels "empty": hammer point 0
    tex = PointTexture(8, N, activation=act, init_method='rand').to(dev())
    ref_t = tex.texture_.detach().cpu().clone().requires_grad_(True)
    want = unet_ref.point_texture(ref_t, ids, act)
    got = tex(ids.to(dev()))
    assert tuple(got.shape) == (2, 8, 40, 56)
    tol = 0 if act == "none" else 1e-6
    assert float((got.detach().cpu() - want.detach()).abs().max()) <= tol
    w = torch.rand(want.shape, generator=g)
    (want * w).sum().backward()
    (got * w.to(dev())).sum().backward()
    gerr = (tex.texture_.grad.cpu() - ref_t.grad).abs().max() / ref_t.grad.abs().max()
    assert float(gerr) < 1e-5, float(gerr)                # descriptor-gradient tolerance (fp32 atomics order)
    with torch.no_grad():
        got2 = tex(ids.to(dev()))                          # inference path: cached shadow + fused activation
    assert float((got2.cpu() - want.detach()).abs().max()) <= 1e-6
