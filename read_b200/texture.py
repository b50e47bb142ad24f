"""Drop-in for ``READ.models.texture.PointTexture`` (READ/models/texture.py:14-70).

Same constructor, same parameter (``texture_`` [1,C,N] f32, so checkpoints load unchanged), same forward
contract (ids [B,1|3,h,w] float -> [B,C,h,w] f32).  The gather and its backward (scatter-add into
``texture_.grad``) are hand-written CUDA kernels reading a point-major [N,C] shadow of the parameter.
"""
import torch
import torch.nn as nn

from . import ops
from . import _lib as L


class Texture(nn.Module):
    """Interface of the reference's texture modules (READ/models/texture.py:6-11): a regulariser that defaults to zero and a
    ``null_grad`` every concrete texture must provide (train.py calls it when a dataset is unloaded)."""

    def null_grad(self):
        raise NotImplementedError(f"{type(self).__name__} does not implement null_grad()")

    def reg_loss(self):
        return 0.


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, texture_, ids):
        tex_nd = ops.texture_to_point_major(texture_)
        ctx.save_for_backward(ids)
        ctx.n = texture_.shape[-1]
        return ops.gather_from_index(tex_nd, ids, L.FEAT_NCHW_F32)

    @staticmethod
    def backward(ctx, grad_out):
        (ids,) = ctx.saved_tensors
        g_nd = ops.gather_backward(grad_out, ids, ctx.n)        # [N,C]
        return ops.texture_to_channel_major(g_nd), None         # [1,C,N]


_INITIALISERS = {'zeros': torch.zeros, 'rand': torch.rand}


class PointTexture(Texture):
    """Per-point descriptors.  Constructor contract of READ/models/texture.py:14-35: ``texture_`` is a float32 Parameter of shape
    [1, num_channels, size] (channel-major, the checkpoint layout), filled by ``init_method`` ('zeros' | 'rand'), or taken from a
    pickled ``{'texture': module}`` checkpoint; ``activation`` in {'none', 'sigmoid', 'tanh'} is applied to the samples."""

    def __init__(self, num_channels, size, activation='none', checkpoint=None, init_method='zeros', reg_weight=0.):
        super().__init__()
        assert isinstance(size, int), 'size must be int'
        if checkpoint:
            descriptors = torch.load(checkpoint, map_location='cpu')['texture'].texture_
        else:
            make = _INITIALISERS.get(init_method)
            if make is None:
                raise ValueError(init_method)
            descriptors = nn.Parameter(make((1, num_channels, size), dtype=torch.float32))
        self.texture_ = descriptors
        self.activation, self.reg_weight = activation, reg_weight
        self._shadow = self._shadow_key = None          # point-major copy for the gather kernels, see point_major()
        self._sparse, self._sparse_requested = None, False     # read_b200.train: sparse gradient accumulator

    def null_grad(self):
        self.texture_.grad = None

    def reg_loss(self):
        """L2 regulariser of texture.py:40-41: reg_weight * mean(texture^2)."""
        return self.reg_weight * self.texture_.square().mean()

    def point_major(self):
        """[N,C] shadow of ``texture_`` on its device, refreshed whenever the parameter changes."""
        t = self.texture_
        key = (t.data_ptr(), t._version, t.device)
        if self._shadow_key != key:
            self._shadow = ops.texture_to_point_major(t.detach())
            self._shadow_key = key
        return self._shadow

    def forward(self, inputs):
        if isinstance(inputs, dict):
            ids = None
            for f, x in inputs.items():
                if 'uv' in f:
                    ids = x[:, 0]
            assert ids is not None, 'Input format does not have uv'
        else:
            ids = inputs[:, 0]                                   # BxHxW
        if not self.texture_.is_cuda:
            raise RuntimeError("read_b200.PointTexture: texture must be on a CUDA device (no CPU fallback)")
        ids = ids.to(self.texture_.device, torch.float32).contiguous()
        if torch.is_grad_enabled() and self.texture_.requires_grad:
            if getattr(self, '_sparse_requested', False):
                # training with read_b200.train.SparseRMSprop: the backward scatter-adds into a persistent point-major accumulator
                # and flags the touched points; texture_.grad is never materialised
                from . import train
                train.enable_sparse_grad(self)
                sample = train._GatherSparse.apply(self.texture_, ids, self)
            else:
                sample = _Gather.apply(self.texture_, ids)
            if self.activation == 'sigmoid':
                return torch.sigmoid(sample)
            if self.activation == 'tanh':
                return torch.tanh(sample)
            return sample
        return ops.gather_from_index(self.point_major(), ids, L.FEAT_NCHW_F32, self.activation)
