"""-m gpu: the descriptor side of the training step (read_b200/train.py, csrc/train.cu) against torch autograd + the dense
torch.optim.RMSprop the reference uses (READ/models/texture.py:55-63, READ/pipelines/ogl.py:16,97-102)."""
import numpy as np
import pytest
import torch

from gpu_util import dev
from read_b200 import train, _lib as L
from read_b200.texture import PointTexture

pytestmark = pytest.mark.gpu


def _ids(gen, B, h, w, N, frac_empty=0.3, pool=None):
    """Index maps like the rasterizer's: float ids, 0 = empty."""
    src = pool if pool is not None else torch.arange(1, N)
    ids = src[torch.randint(0, len(src), (B, 1, h, w), generator=gen)].float()
    ids[torch.rand((B, 1, h, w), generator=gen) < frac_empty] = 0.
    return ids


def test_sparse_backward_equals_autograd_of_the_oracle():
    from oracle import unet_ref
    N, D = 5000, 8
    gen = torch.Generator().manual_seed(0)
    tex = PointTexture(D, N, init_method='rand').to(dev())
    ids = _ids(gen, 3, 24, 40, N)
    up = torch.randn((3, D, 24, 40), generator=gen)
    ref_param = tex.texture_.detach().cpu().clone().requires_grad_(True)
    (unet_ref.point_texture(ref_param, ids) * up).sum().backward()
    train.enable_sparse_grad(tex)
    out = tex(ids.to(dev()))
    (out * up.to(dev())).sum().backward()
    assert tex.texture_.grad is None                                     # nothing dense was materialised
    st = tex._sparse
    got = st.grad.cpu().t()                                              # [D, N]
    want = ref_param.grad[0]
    assert float((got - want).abs().max()) < 1e-4 * float(want.abs().max())
    touched = st.touched.cpu().bool()
    assert torch.equal(touched, (torch.bincount(ids.long().flatten(), minlength=N) > 0))
    assert train.touched_count(tex) == int(touched.sum())


def test_sparse_rmsprop_equals_dense_torch_rmsprop_with_lazy_decay():
    """8 steps; every step touches a different random subset (so most points skip most steps): parameters and the materialised
    square_avg must equal torch.optim.RMSprop run densely on the same gradients."""
    N, D = 20000, 8
    gen = torch.Generator().manual_seed(1)
    tex = PointTexture(D, N, init_method='rand').to(dev())
    ref = torch.nn.Parameter(tex.texture_.detach().clone())
    opt_ref = torch.optim.RMSprop([ref], lr=0.1)
    opt = train.SparseRMSprop(tex, lr=0.1)
    for step in range(8):
        pool = torch.randperm(N - 1, generator=gen)[: 300 + 700 * (step % 3)] + 1
        ids = _ids(gen, 2, 32, 32, N, pool=pool).to(dev())
        if step == 4:
            opt.param_groups[0]['lr'] = 0.05                             # the pipeline rescales lr through param_groups
            opt_ref.param_groups[0]['lr'] = 0.05
        up = torch.randn((2, D, 32, 32), generator=gen).to(dev())
        # dense reference: autograd through index_select on the reference parameter
        opt_ref.zero_grad()
        idx = ids[:, 0].long().reshape(-1)
        smp = torch.index_select(ref[0], 1, idx).view(D, 2, 32, 32).permute(1, 0, 2, 3)
        (smp * up).sum().backward()
        opt_ref.step()
        opt.zero_grad()
        (tex(ids) * up).sum().backward()
        opt.step()
        assert train.touched_count(tex) == 0                             # flags cleared by the step
        assert float(tex._sparse.grad.abs().max()) == 0.0                # and the touched gradient rows
    torch.cuda.synchronize()
    err = float((tex.texture_.detach() - ref.detach()).abs().max())
    assert err < 2e-5, err
    sq = opt.dense_square_avg(tex)
    sq_ref = opt_ref.state[ref]['square_avg']
    assert float((sq - sq_ref).abs().max()) < 1e-5 * float(sq_ref.abs().max()) + 1e-9
    # the point-major shadow the gather kernels read was kept in sync without a dense re-transposition
    assert torch.equal(tex.point_major(), tex.texture_[0].t().contiguous())
    # state_dict round trip in torch.optim.RMSprop's layout
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and tuple(sd["state"][0]["square_avg"].shape) == (1, D, N)
    opt2 = train.SparseRMSprop(tex, lr=0.1)
    tex(ids)                                                             # creates the sparse state
    opt2.load_state_dict(sd)
    assert float((opt2.dense_square_avg(tex) - sq).abs().max()) < 1e-7


def test_compact_and_scatter_pairs_round_trip():
    lib = L.load()
    N, D = 10000, 8
    gen = torch.Generator().manual_seed(2)
    grad = torch.zeros((N, D), device=dev())
    touched = torch.zeros(N, dtype=torch.uint8, device=dev())
    sel = torch.randperm(N, generator=gen)[:777].to(dev())
    grad[sel] = torch.randn((777, D), generator=gen).to(dev())
    touched[sel] = 1
    cnt = torch.zeros(1, dtype=torch.int32, device=dev())
    ids = torch.empty(1024, dtype=torch.int32, device=dev())
    vals = torch.empty((1024, D), device=dev())
    L.check(lib.read_compact_touched(grad.data_ptr(), touched.data_ptr(), N, D, cnt.data_ptr(), 1024, ids.data_ptr(), vals.data_ptr(), L.stream_ptr()))
    n = int(cnt.item())
    assert n == 777 and torch.equal(torch.sort(ids[:n].long()).values, torch.sort(sel).values)
    g2 = torch.zeros_like(grad)
    t2 = torch.zeros_like(touched)
    L.check(lib.read_scatter_pairs(ids.data_ptr(), vals.data_ptr(), n, D, N, g2.data_ptr(), t2.data_ptr(), L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(g2, grad) and torch.equal(t2, touched)


def test_exchange_is_identity_without_a_process_group():
    tex = PointTexture(8, 1000, init_method='rand').to(dev())
    train.enable_sparse_grad(tex)
    ids = torch.randint(1, 1000, (1, 1, 8, 8)).float().to(dev())
    tex(ids).sum().backward()
    before = tex._sparse.grad.clone()
    assert train.exchange_sparse_grads(tex) == train.touched_count(tex)
    assert torch.equal(tex._sparse.grad, before)
