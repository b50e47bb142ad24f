"""CPU test of the strip-parallel plan (read_b200/strips.py + engine.StripEngine's halo bookkeeping): the net is executed on the
crops of R emulated ranks with torch CPU ops - zero padding at every crop edge, exactly what the CUDA kernels do - and halo rows
are copied between neighbouring crops wherever (and only where) the engine's plan inserts an exchange.  The stitched strips must
equal the full-frame oracle.  The graph walked here IS engine.UNetEngine._build (same methods); only `_conv` is re-expressed in
torch."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from read_b200 import strips, synth
from read_b200.engine import StripEngine, _Layer


class MR:
    """One activation tensor on all emulated ranks: per-rank NCHW crops; ``shape`` mimics the NHWC shape of an interior rank."""

    def __init__(self, crops, level, S, W):
        self.crops, self.level = crops, level
        self.shape = (1, (S >> level) + 2 * strips.halo_rows(level), W >> level, crops[0].shape[1])


def _resample(x, mode, f):
    if mode == "id":
        return x
    if mode == "down":
        return x[:, :, ::f, ::f]
    if mode == "up":
        return x.repeat_interleave(f, 2).repeat_interleave(f, 3)
    return F.interpolate(x, scale_factor=4, mode="bilinear", align_corners=False)


class SimStrip(StripEngine):
    def __init__(self, sd, H, W, world):
        self.rank, self.world, self.group = 0, world, None
        self.H_full, self.S = H, H // world
        self._strips = strips
        self._meta, self._exch = {}, []
        self.h_top = self.h_bot = strips.HALO0
        self.B, self.W, self.H = 1, W, self.S + 2 * strips.HALO0           # nominal interior rank
        self.bf16, self.conv_impl, self.base, self.num_res = True, "auto", 32, 4
        self.sd = {k: v.float() for k, v in sd.items() if v.dtype.is_floating_point}
        self.layers, self.ops, self.n_exchanged_rows = [], [], 0

    def crop(self, full, level):
        """full NCHW tensor at ``level`` -> per-rank crops."""
        out = []
        for r in range(self.world):
            a, n, _, _ = strips.strip_rows(self.H_full, self.world, r, level)
            out.append(full[:, :, a:a + n].clone())
        return MR(out, level, self.S, self.W)

    def stitch(self, mr):
        rows = []
        for r in range(self.world):
            _, n, top, bot = strips.strip_rows(self.H_full, self.world, r, mr.level)
            rows.append(mr.crops[r][:, :, top:n - bot])
        return torch.cat(rows, 2)

    def _exchange(self, t):
        super()._exchange(t)                     # bookkeeping (validity := h) + the op record
        h = strips.halo_rows(t.level)
        for r in range(self.world - 1):
            up, dn = t.crops[r], t.crops[r + 1]
            _, n_up, top_up, bot_up = strips.strip_rows(self.H_full, self.world, r, t.level)
            _, n_dn, top_dn, _ = strips.strip_rows(self.H_full, self.world, r + 1, t.level)
            dn[:, :, :h] = up[:, :, n_up - bot_up - h:n_up - bot_up]          # my last interior rows -> lower rank's top halo
            up[:, :, n_up - bot_up:] = dn[:, :, top_dn:top_dn + h]             # lower rank's first interior rows -> my bottom halo
        self.n_exchanged_rows += h

    def _merge_src(self, t):
        if self._v(t)[1] < 1:
            self._exchange(t)
        out = MR([_resample(c, "bil4", 4) for c in t.crops], t.level - 2, self.S, self.W)
        m = self._v(t)
        self._meta[id(out)] = [m[0] - 2, strips.src_validity(m[1], "bil4", 4, strips.halo_rows(m[0] - 2))]
        return (out, "id", 1)

    def _conv(self, prefix, srcs, cout, k, stride, elu, residual=None, out2_mul=None, final=False, raw=False, addin=None,
              cin_slice=None, name=None):
        self._before_conv(srcs, k, stride, residual, out2_mul, addin)
        sd, p = self.sd, int((k - 1) / 2)
        wf, wm = sd[prefix + ".block.conv_f.weight"], sd[prefix + ".block.conv_m.weight"]
        if cin_slice is not None:
            wf, wm = wf[:, cin_slice[0]:cin_slice[1]], wm[:, cin_slice[0]:cin_slice[1]]
        bf, bm = sd[prefix + ".block.conv_f.bias"], sd[prefix + ".block.conv_m.bias"]
        n = prefix + ".block.norm."
        scale = sd[n + "weight"] / torch.sqrt(sd[n + "running_var"] + 1e-5)
        shift = sd[n + "bias"] - sd[n + "running_mean"] * scale
        outs, outs2 = [], []
        for r in range(self.world):
            x = torch.cat([_resample(t.crops[r], mode, f) for (t, mode, f) in srcs], 1)
            f_ = F.conv2d(x, wf, None, stride=stride, padding=p)
            m_ = F.conv2d(x, wm, None, stride=stride, padding=p)
            if addin is not None:
                a = _resample(addin.crops[r], "up", 2)[:, :, :f_.shape[2], :f_.shape[3]]
                f_, m_ = f_ + a[:, :cout], m_ + a[:, cout:]
            if raw:
                outs.append(torch.cat([f_, m_], 1))
                continue
            f_ = f_ + bf[None, :, None, None]
            if elu:
                f_ = F.elu(f_)
            y = f_ * torch.sigmoid(m_ + bm[None, :, None, None]) * scale[None, :, None, None] + shift[None, :, None, None]
            if residual is not None:
                y = y + residual.crops[r]
            outs.append(y)
            if out2_mul is not None:
                outs2.append(y * out2_mul.crops[r])
        t0, mode0, f0 = srcs[0]
        lg = {1: 0, 2: 1, 4: 2, 8: 3}[f0]
        lvl = t0.level + (lg if mode0 == "down" else (-lg if mode0 == "up" else (-2 if mode0 == "bil4" else 0)))
        lvl += 1 if stride == 2 else 0
        out = MR(outs, lvl, self.S, self.W)
        out2 = MR(outs2, lvl, self.S, self.W) if out2_mul is not None else None
        op = _Layer()
        op.name, op.kind, op.plan = name or prefix, "conv", None
        self.ops.append(op)
        self._after_conv(srcs, k, stride, out, out2, residual, out2_mul, addin, final)
        return (out, out2) if out2_mul is not None else out

    def run_full(self, feats):
        """feats: 4 full-frame NCHW tensors -> stitched RGB [1,3,H,W]."""
        self.inputs = [self.crop(f, l) for l, f in enumerate(feats)]
        return self.stitch(self._build_graph(self.inputs))


@pytest.mark.parametrize("world,H,W", [(2, 64, 32), (4, 128, 16), (2, 96, 48)])
def test_strip_plan_is_exact(world, H, W):
    from oracle import unet_ref
    sd = synth.synth_state_dict(synth.SEED)
    g = torch.Generator().manual_seed(world * 100 + H)
    feats = [torch.rand((1, 8, H >> l, W >> l), generator=g) for l in range(4)]
    with torch.no_grad():
        want = unet_ref.unet_forward(sd, feats)
        sim = SimStrip(sd, H, W, world)
        got = sim.run_full(feats)
    n_ex = sum(1 for op in sim.ops if op.kind == "halo")
    n_conv = sum(1 for op in sim.ops if op.kind == "conv")
    assert n_conv == 102 and 2 <= n_ex < n_conv, (n_conv, n_ex)        # far fewer exchanges than layers
    err = float((got - want).abs().max())
    assert err < 5e-5, err
    # the plan is needed: without exchanges the strips disagree with the full frame
    sim2 = SimStrip(sd, H, W, world)
    sim2._exchange = lambda t: StripEngine._exchange(sim2, t)            # bookkeeping only, no data movement
    with torch.no_grad():
        bad = sim2.run_full(feats)
    assert float((bad - want).abs().max()) > 20 * max(err, 1e-7), (float((bad - want).abs().max()), err)


def test_validity_rules():
    assert strips.halo_rows(0) == 16 and strips.halo_rows(3) == 2 and strips.halo_rows(4) == 1
    assert strips.conv_out_validity(2, 3, 1) == 1 and strips.conv_out_validity(0, 3, 1) == 0
    assert strips.conv_out_validity(16, 3, 2) == 7 and strips.conv_out_validity(2, 4, 2) == 0
    assert strips.src_validity(5, "down", 2, 8) == 2 and strips.src_validity(2, "up", 4, 4) == 4
    assert strips.src_validity(1, "bil4", 4, 16) == 2 and strips.src_validity(0, "bil4", 4, 16) == 0
    a, n, top, bot = strips.strip_rows(1088, 2, 1, 0)
    assert (a, n, top, bot) == (544 - 16, 544 + 16, 16, 0)
    a, n, top, bot = strips.strip_rows(1088, 4, 1, 3)
    assert (a, n, top, bot) == (34 - 2, 34 + 4, 2, 2)
    with pytest.raises(AssertionError):
        strips.strip_rows(1088, 8, 0)
