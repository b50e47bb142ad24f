"""Helpers for the -m gpu parity tests (they call the product path through the C-ABI library)."""
import numpy as np
import torch

from read_b200 import ops, synth


def dev():
    return torch.device("cuda", 0)


def render_gpu(xyz, total_m, W, H, L, id_base=0):
    """numpy in -> pyramid rendered by the CUDA library -> (index_l, depth_l) numpy lists [B,h,w]."""
    d = dev()
    x = torch.from_numpy(np.ascontiguousarray(xyz, np.float32)).to(d)
    m = torch.from_numpy(np.ascontiguousarray(total_m, np.float32)).to(d)
    pyr = ops.Pyramid(m.shape[0], W, H, L, d)
    pyr.clear()
    ops.raster_project(pyr, x, m, id_base=id_base)
    out = [ops.zbuf_resolve(pyr, l) for l in range(L)]
    torch.cuda.synchronize()
    return [o[0].cpu().numpy() for o in out], [o[1].cpu().numpy() for o in out], pyr


def scene_and_cams(n, W, H, ts, depth=60.0, seed=1):
    xyz = synth.street_scene(n, depth=depth, seed=seed)
    proj, view = synth.camera_batch(W, H, ts)
    return xyz, synth.total_matrix(proj, view)


def psnr(a, b, peak=1.0):
    mse = float(((a - b) ** 2).mean())
    return 99.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)
