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


@pytest.mark.parametrize("layout", [L.FEAT_NHWC_BF16, L.FEAT_NHWC_F32])
def test_fused_pyramid_resolve_equals_separate_kernels(layout):
    """derive + 4 gathers + clear in ONE kernel must give the same pyramid and the same feature maps, bit for bit."""
    from gpu_util import scene_and_cams
    W, H, B = 96, 64, 3
    xyz, M = scene_and_cams(60_000, W, H, [0, 3, 8])
    d = dev()
    x, m = torch.from_numpy(xyz).to(d), torch.from_numpy(M).to(d)
    nd = torch.rand(60_000, 8, device=d)
    ref = ops.Pyramid(B, W, H, 4, d)
    ref.clear()
    ops.raster_project(ref, x, m)                                   # separate path (derive kernels)
    want = [ops.gather_from_zbuf(nd, ref, l, layout) for l in range(4)]
    pyr = ops.Pyramid(B, W, H, 4, d)
    pyr.clear()
    ops.raster_project(pyr, x, m, derive=False)
    dt = torch.bfloat16 if layout == L.FEAT_NHWC_BF16 else torch.float32
    outs = [torch.full((B, H >> l, W >> l, 8), float("nan"), dtype=dt, device=d) for l in range(4)]
    ops.pyramid_resolve_gather(nd, pyr, outs, layout, reset_level0=False)
    assert torch.equal(pyr.buf, ref.buf)
    for l in range(4):
        assert torch.equal(outs[l], want[l]), l
    # view sub-range + reset: only view 1, level 0 of that view cleared afterwards
    outs1 = [torch.empty((1, H >> l, W >> l, 8), dtype=dt, device=d) for l in range(4)]
    ops.pyramid_resolve_gather(nd, pyr, outs1, layout, view0=1, nviews=1, reset_level0=True)
    for l in range(4):
        assert torch.equal(outs1[l][0], want[l][1])
    lvl0 = pyr.level(0).view(B, H, W)
    assert bool((lvl0[1] == 0x7FFFFFFFFFFFFFFF).all()) and torch.equal(lvl0[0], ref.level(0).view(B, H, W)[0])


def test_fused_render_twice_is_stable(synth_sd=None):
    """NetAndTexture.render keeps level 0 clean between frames: the same camera twice gives identical frames."""
    from read_b200 import synth
    from read_b200.unet import UNet
    from read_b200.compose import NetAndTexture
    from gpu_util import scene_and_cams
    xyz, M = scene_and_cams(40_000, 64, 64, [1])
    net = UNet()
    net.load_state_dict(synth.synth_state_dict(synth.SEED), strict=True)
    tex = PointTexture(8, 40_000, init_method='rand')
    model = NetAndTexture(net, {0: tex}, 1)
    model.load_textures(0)
    model.cuda().eval()
    x, m = torch.from_numpy(xyz).cuda(), torch.from_numpy(M).cuda()
    a = model.render(x, m, 64, 64).clone()
    b = model.render(x, m, 64, 64).clone()
    c, maps = model.render(x, m, 64, 64, want_maps=True)         # non-fused path
    assert torch.equal(a, b) and torch.equal(a, c)
