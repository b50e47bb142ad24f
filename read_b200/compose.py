"""Drop-in for ``READ.models.compose.NetAndTexture`` (READ/models/compose.py:84-181) and ``ModelAndLoss``
(compose.py:12-32).

Same constructor, attributes (``ss``, ``temporal_average``, ``last_input``) and texture management
(``load_textures`` / ``unload_textures`` / ``reg_loss``).  ``forward`` keeps the reference's input contract
(dict of index maps keyed by format string + 'id') and output ([B,3,H,W]); batch items that share a texture are
run as ONE batched net pass (equivalent under eval-mode BatchNorm, SURVEY.md §8a "Batching").

Extra fast path (not in the reference): ``render(xyz, total_m, W, H)`` goes points -> packed z-buffer pyramid ->
feature pyramid -> net without ever materialising the float index maps.
"""
import torch
import torch.nn as nn

from . import ops
from . import _lib as L


class NetAndTexture(nn.Module):
    def __init__(self, net, textures, supersampling=1, temporal_average=False):
        super().__init__()
        self.net = net
        self.ss = supersampling
        try:
            textures = dict(textures)
        except TypeError:
            textures = {0: textures}
        self._textures = {k: v.cpu() for k, v in textures.items()}
        self._loaded_textures = []
        self.last_input = None
        self.temporal_average = temporal_average

    def load_textures(self, texture_ids):
        if torch.is_tensor(texture_ids):
            texture_ids = texture_ids.cpu().tolist()
        elif isinstance(texture_ids, int):
            texture_ids = [texture_ids]
        for tid in texture_ids:
            self._modules[str(tid)] = self._textures[tid]
        self._loaded_textures = texture_ids

    def unload_textures(self):
        for tid in self._loaded_textures:
            self._modules[str(tid)].cpu()
            del self._modules[str(tid)]

    def reg_loss(self):
        loss = 0
        for tid in self._loaded_textures:
            loss += self._modules[str(tid)].reg_loss()
        return loss

    def _sample_item(self, texture, item):
        """compose.py:143-165 for one batch item (dict of [1,C,h,w])."""
        keys = list(item)
        assert 'uv' in keys[0], 'first input must be uv'
        j, ms = 0, []
        while j < len(keys):
            assert 'uv' in keys[j]
            tex_sample = texture(item[keys[j]])
            j += 1
            extra = []
            while j < len(keys) and 'uv' not in keys[j]:
                extra.append(item[keys[j]])
                j += 1
            cat = torch.cat(extra + [tex_sample], 1) if extra else tex_sample
            if self.ss > 1:
                cat = nn.functional.interpolate(cat, scale_factor=1. / self.ss, mode='bilinear')
            ms.append(cat)
        return ms

    def forward(self, inputs, **kwargs):
        inputs = dict(inputs)
        texture_ids = inputs.pop('id')
        if torch.is_tensor(texture_ids):
            texture_ids = texture_ids.tolist()
        elif isinstance(texture_ids, int):
            texture_ids = [texture_ids]
        texture_ids = [int(t) for t in texture_ids]

        batched = (not self.temporal_average) and len(set(texture_ids)) == 1 and len(texture_ids) > 1
        if batched:
            texture = self._modules[str(texture_ids[0])]
            input_multiscale = self._sample_item(texture, inputs)
            out = self.net(*input_multiscale, **kwargs)
        else:
            outs = []
            for i, tid in enumerate(texture_ids):                 # per item in batch (compose.py:136)
                item = {k: v[i][None] for k, v in inputs.items()}
                texture = self._modules[str(tid)]
                input_multiscale = self._sample_item(texture, item)
                if self.temporal_average:
                    if self.last_input is not None:
                        for j in range(len(input_multiscale)):
                            input_multiscale[j] = (input_multiscale[j] + self.last_input[j]) / 2
                    self.last_input = list(input_multiscale)
                outs.append(self.net(*input_multiscale, **kwargs))
            out = torch.cat(outs, 0)
        if kwargs.get('return_input'):
            return out, input_multiscale
        return out

    # ------------------------------------------------------------------ fused fast path
    @torch.no_grad()
    def render(self, xyz, total_m, W, H, texture_id=0, n_levels=4, want_maps=False):
        """points [N,3] (cuda f32) or an ``ops.SortedPoints`` store + total_m [B,4,4] (cuda f32) -> RGB [B,3,H,W] f32, all on
        device, one pass over the cloud.  A sorted store serves single-view frames with nested levels; the result is
        bit-identical to rendering the unsorted cloud (the z-buffer is a min over (depth | original id) keys)."""
        store = xyz if isinstance(xyz, ops.SortedPoints) else None
        if store is not None:
            xyz = store.pts4
        L.require_device()
        texture = self._modules[str(texture_id)]
        B = total_m.shape[0]
        eng = self.net.engine(B, H, W, xyz.device)
        pyr = getattr(self, "_pyr", None)
        if pyr is None or (pyr.B, pyr.W, pyr.H, pyr.L) != (B, W, H, n_levels) or pyr.buf.device != xyz.device:
            pyr = self._pyr = ops.Pyramid(B, W, H, n_levels, xyz.device)
        layout = L.FEAT_NHWC_BF16 if eng.bf16 else L.FEAT_NHWC_F32
        tex = texture.point_major()
        if (not want_maps) and texture.activation == 'none' and ops.fused_resolve_supported(pyr, tex.shape[1]):
            # 2 launches: project all points into level 0, then ONE kernel derives levels 1..3, gathers the four
            # feature maps and leaves level 0 cleared for the next frame
            if not getattr(pyr, "level0_clean", False):
                pyr.clear()
            if store is not None:
                if B != 1 or pyr.direct_mask != 1:
                    raise RuntimeError("a SortedPoints store renders one view with nested levels; pass the [N,3] cloud otherwise")
                ops.raster_project_sorted(pyr, store, total_m)
            else:
                ops.raster_project(pyr, xyz, total_m, derive=False)
            ops.pyramid_resolve_gather(tex, pyr, eng.inputs, layout, reset_level0=True)
            pyr.level0_clean = True
        else:
            pyr.clear()
            pyr.level0_clean = False
            if store is not None:
                if B != 1 or pyr.direct_mask != 1:
                    raise RuntimeError("a SortedPoints store renders one view with nested levels; pass the [N,3] cloud otherwise")
                ops.raster_project_sorted(pyr, store, total_m)
                ops.raster_derive(pyr)
            else:
                ops.raster_project(pyr, xyz, total_m)
            for l in range(4):
                ops.gather_from_zbuf(tex, pyr, l, layout, texture.activation, out=eng.inputs[l])
        out = eng.run()
        if want_maps:
            return out, [ops.zbuf_resolve(pyr, l) for l in range(n_levels)]
        return out


class ModelAndLoss(nn.Module):
    """compose.py:12-32: wraps model + criterion so DataParallel scatters both."""

    def __init__(self, model, loss, use_mask=False):
        super().__init__()
        self.model = model
        self.loss = loss
        self.use_mask = use_mask

    def forward(self, *args, **kwargs):
        input = args[:-1]
        target = args[-1]
        if not isinstance(input, (tuple, list)):
            input = [input]
        output = self.model(*input, **kwargs)
        if self.use_mask and 'mask' in kwargs and kwargs['mask'] is not None:
            loss = self.loss(output * kwargs['mask'], target)
        else:
            loss = self.loss(output, target)
        return output, loss
