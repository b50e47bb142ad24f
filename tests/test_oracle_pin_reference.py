"""Pin the oracle's torch restatement to the reference's real modules (only where /root/reference exists,
i.e. in the build container; skipped on the GPU box)."""
import sys
import types

import numpy as np
import pytest
import torch

from conftest import HAVE_REFERENCE
from read_b200 import synth

pytestmark = pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference not present")


def _ref():
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    from READ.models.unet import UNet
    from READ.models.texture import PointTexture
    from READ.models.compose import NetAndTexture
    return UNet, PointTexture, NetAndTexture


def test_state_dict_keys_identical_to_reference(synth_sd):
    UNet, _, _ = _ref()
    ref_sd = UNet().state_dict()
    assert set(ref_sd) == set(synth_sd)
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(synth_sd[k].shape), k
    from read_b200.unet import UNet as OurUNet
    ours = OurUNet().state_dict()
    assert list(sorted(ours)) == list(sorted(ref_sd))
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(ours[k].shape) and v.dtype == ours[k].dtype, k
    OurUNet().load_state_dict(ref_sd, strict=True)


def test_unet_oracle_equals_reference_module(synth_sd):
    from oracle import unet_ref
    UNet, _, _ = _ref()
    net = UNet()
    net.load_state_dict(synth_sd, strict=True)
    net.eval()
    g = torch.Generator().manual_seed(7)
    H, W = 32, 48
    xs = [torch.rand((2, 8, H >> l, W >> l), generator=g) for l in range(5)]
    with torch.no_grad():
        want = net(*xs)
        got = unet_ref.unet_forward(synth_sd, xs)
    assert float((want - got).abs().max()) < 2e-5


def test_gather_oracle_equals_reference_module():
    from oracle import unet_ref
    _, PointTexture, _ = _ref()
    g = torch.Generator().manual_seed(3)
    tex = PointTexture(8, 500, init_method='rand')
    ids = torch.randint(0, 500, (3, 1, 9, 7), generator=g).float()
    with torch.no_grad():
        want = tex(ids)
    got = unet_ref.point_texture(tex.texture_.detach(), ids)
    assert torch.equal(want, got)


def test_training_forward_of_our_unet_equals_reference(synth_sd):
    """The library (autograd) path of read_b200.unet.UNet must agree with the reference on CPU."""
    UNet, _, _ = _ref()
    from read_b200.unet import UNet as OurUNet
    ref, ours = UNet(), OurUNet()
    ref.load_state_dict(synth_sd)
    ours.load_state_dict(synth_sd)
    ref.eval()
    ours.eval()
    g = torch.Generator().manual_seed(11)
    xs = [torch.rand((1, 8, 32 >> l, 32 >> l), generator=g) for l in range(4)]
    want = ref(*xs)
    got = ours(*xs)                      # grad enabled -> torch path
    assert float((want - got).abs().max()) < 2e-5
    got.sum().backward()
    assert ours.get_submodule("feat_extract.0").block["conv_f"].weight.grad is not None
