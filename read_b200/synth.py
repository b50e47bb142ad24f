"""Seeded synthetic scenes, cameras and weights for the benchmark configs.

The reference ships no data (scenes/weights are a zenodo download, README.md:59), so every config of
BASELINE.json runs on these generators.  ``get_proj_matrix`` restates READ/gl/utils.py:123-150.
"""
import numpy as np

SEED = 2019     # reference default seed, train.py:457


def get_proj_matrix(K, image_size, znear=.01, zfar=1000.):
    """GL projection from pinhole intrinsics, returned transposed exactly like READ/gl/utils.py:123-150."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    width, height = image_size
    m = np.zeros((4, 4))
    m[0][0] = 2.0 * fx / width
    m[1][1] = 2.0 * fy / height
    m[2][0] = 1.0 - 2.0 * cx / width
    m[2][1] = 2.0 * cy / height - 1.0
    m[2][2] = (zfar + znear) / (znear - zfar)
    m[2][3] = -1.0
    m[3][2] = 2.0 * zfar * znear / (znear - zfar)
    return m.T


def street_scene(n_points, depth=250.0, seed=SEED):
    """Camera-frame GL street (x right, y up, -z forward): 40% ground, 40% two facades, 20% box clutter.
    Generation order is kept (no shuffle), float32 [n,3]."""
    rng = np.random.default_rng(seed)
    n_ground = int(0.4 * n_points)
    n_fac = int(0.4 * n_points)
    n_clut = n_points - n_ground - n_fac
    g = np.empty((n_ground, 3), np.float32)
    g[:, 0] = rng.uniform(-15, 15, n_ground)
    g[:, 1] = -1.6 + rng.normal(0, 0.02, n_ground)
    g[:, 2] = rng.uniform(-depth, 0, n_ground)
    f = np.empty((n_fac, 3), np.float32)
    side = np.where(np.arange(n_fac) < n_fac // 2, -15.0, 15.0)
    f[:, 0] = side + rng.normal(0, 0.05, n_fac)
    f[:, 1] = rng.uniform(-1.6, 12, n_fac)
    f[:, 2] = rng.uniform(-depth, 0, n_fac)
    n_boxes = 2000
    per = np.full(n_boxes, n_clut // n_boxes)
    per[: n_clut - per.sum()] += 1
    centers = np.stack([rng.uniform(-13, 13, n_boxes), np.full(n_boxes, -1.6), rng.uniform(-depth, 0, n_boxes)], 1)
    sizes = rng.uniform(0.3, 4.0, (n_boxes, 3))
    box_id = np.repeat(np.arange(n_boxes), per)
    u = rng.uniform(-0.5, 0.5, (n_clut, 3))
    face = rng.integers(0, 3, n_clut)
    sign = rng.integers(0, 2, n_clut) * 1.0 - 0.5
    u[np.arange(n_clut), face] = sign                       # snap one coordinate onto a box face
    c = centers[box_id] + u * sizes[box_id]
    c[:, 1] += sizes[box_id, 1] * 0.5                       # boxes sit on the ground
    return np.concatenate([g, f, c.astype(np.float32)], 0).astype(np.float32)


def intrinsics(W, H):
    return np.array([[0.8 * W, 0, W / 2.0], [0, 0.8 * W, H / 2.0], [0, 0, 1.0]])


def camera_pose(t):
    """Camera-to-world 4x4 (GL convention) of trajectory step t: eye (0,0,-0.5t), yaw 3deg*sin(t/8)."""
    yaw = np.deg2rad(3.0) * np.sin(t / 8.0)
    c, s = np.cos(yaw), np.sin(yaw)
    m = np.eye(4)
    m[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    m[:3, 3] = [0.0, 0.0, -0.5 * t]
    return m


def camera_batch(W, H, ts, znear=0.1, zfar=1000.0):
    """(proj [B,4,4] f32, view [B,4,4] f32) for trajectory steps ``ts`` (znear/zfar: READ/datasets/dynamic.py:111-112)."""
    proj = get_proj_matrix(intrinsics(W, H), (W, H), znear, zfar).astype(np.float32)
    view = np.stack([camera_pose(t) for t in ts]).astype(np.float32)
    return np.repeat(proj[None], len(ts), 0), view


def crop_cameras(W, H, ts, rng, znear=0.1, zfar=1000.0):
    """Training crops (configs/train_example.yaml:30-31: random zoom U(0.7, 2.0) and shift of the view before cropping): one
    projection per crop with the focal length scaled by the zoom and the principal point shifted by up to a quarter frame."""
    projs = []
    for _ in ts:
        K = intrinsics(W, H)
        zoom = rng.uniform(0.7, 2.0)
        K[0, 0] *= zoom
        K[1, 1] *= zoom
        K[0, 2] += rng.uniform(-0.25, 0.25) * W
        K[1, 2] += rng.uniform(-0.25, 0.25) * H
        projs.append(get_proj_matrix(K, (W, H), znear, zfar).astype(np.float32))
    view = np.stack([camera_pose(t) for t in ts]).astype(np.float32)
    return np.stack(projs), view


def total_matrix(proj, view):
    """proj @ inv(view) in float32 with the same numpy call as src/READ/gl/myrender.py:28-30."""
    return (np.asarray(proj, np.float32) @ np.linalg.inv(np.asarray(view, np.float32))).astype(np.float32)


def randomize_bn_(net, seed=SEED):
    """BN running_mean ~ N(0,0.1), running_var ~ U(0.5,1.5), affine ~ U(0.5,1.5)/N(0,0.1) so eval-mode BN is exercised."""
    import torch
    g = torch.Generator().manual_seed(seed)
    for name, buf in net.named_buffers():
        if name.endswith("running_mean"):
            buf.copy_(torch.randn(buf.shape, generator=g) * 0.1)
        elif name.endswith("running_var"):
            buf.copy_(torch.rand(buf.shape, generator=g) + 0.5)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith("norm.weight"):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif name.endswith("norm.bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return net


def synth_state_dict(seed=SEED, base=32, num_res=4):
    """Deterministic UNet weights with the reference's keys/shapes: every tensor is drawn from its own CPU
    generator seeded by crc32(key)^seed (independent of module construction order, so the reference's UNet,
    ours and the oracle can all be loaded with the SAME weights from just a seed).
    Conv weights/biases ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch's default scale); BN affine/running stats
    randomised so the eval-mode affine is exercised."""
    import zlib
    import torch
    from .unet import layer_table
    sd = {}

    def gen(key):
        return torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ seed) & 0x7FFFFFFF)

    for prefix, cin, cout, k, _stride, _elu in layer_table(base, num_res):
        bound = 1.0 / float(cin * k * k) ** 0.5
        for conv in ("conv_f", "conv_m"):
            kw, kb = f"{prefix}.block.{conv}.weight", f"{prefix}.block.{conv}.bias"
            sd[kw] = (torch.rand((cout, cin, k, k), generator=gen(kw)) * 2 - 1) * bound
            sd[kb] = (torch.rand((cout,), generator=gen(kb)) * 2 - 1) * bound
        n = f"{prefix}.block.norm."
        sd[n + "weight"] = torch.rand((cout,), generator=gen(n + "weight")) + 0.5
        sd[n + "bias"] = torch.randn((cout,), generator=gen(n + "bias")) * 0.1
        sd[n + "running_mean"] = torch.randn((cout,), generator=gen(n + "running_mean")) * 0.1
        sd[n + "running_var"] = torch.rand((cout,), generator=gen(n + "running_var")) + 0.5
        sd[n + "num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd


def state_dict_checksum(sd):
    import torch
    return float(sum(v.double().abs().sum() for k, v in sorted(sd.items()) if v.dtype.is_floating_point))
