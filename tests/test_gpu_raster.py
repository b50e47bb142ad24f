"""-m gpu: CUDA rasterizer vs the sequential oracle — bit-exact index and depth (through the C ABI)."""
import numpy as np
import pytest
import torch

from gpu_util import dev, render_gpu, scene_and_cams
from read_b200 import ops, pcpr, synth
from read_b200.myrender import MyRender

pytestmark = pytest.mark.gpu
ID = np.eye(4, dtype=np.float32)[None]


def _check(oracle_mod, xyz, M, W, H, L):
    gi, gd, _ = render_gpu(xyz, M, W, H, L)
    for l, (w, h) in enumerate(oracle_mod.level_sizes(W, H, L)):
        oi, od = oracle_mod.pcpr_forward(xyz, M, w, h)
        np.testing.assert_array_equal(gi[l], oi, err_msg=f"index level {l} ({w}x{h})")
        np.testing.assert_array_equal(gd[l].view(np.uint32), od.view(np.uint32), err_msg=f"depth bits level {l}")


def test_kats_identity_matrix(oracle_mod):
    xyz = np.array([[9, 9, 9], [0.1, 0.1, 0.5], [0.1, 0.1, -0.25], [0.1, 0.1, -0.25], [0.1, 0.1, 0.0],
                    [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, 0]], np.float32)
    _check(oracle_mod, xyz, ID, 10, 10, 1)
    _check(oracle_mod, xyz, ID, 4, 4, 1)
    _check(oracle_mod, xyz, ID, 8, 135, 2)      # 135 -> 67: odd level must use direct atomics


@pytest.mark.parametrize("n,W,H,L,ts", [
    (100_000, 256, 256, 5, [0]),                 # C1
    (50_000, 152, 46, 4, [2]),                   # kitti6 aspect, odd coarse levels (152x46,76x23,38x11,19x5)
    (80_000, 240, 135, 2, [5]),                  # 1080p tail levels 240x135 -> 120x67 (non-nested)
    (30_001, 96, 64, 4, [0, 7, 13]),             # batch of 3 views, N not a multiple of anything
    (1, 16, 16, 3, [0]),
    (1023, 33, 17, 3, [1]),
])
def test_random_scenes_bit_exact(oracle_mod, n, W, H, L, ts):
    xyz, M = scene_and_cams(n, W, H, ts)
    assert all(oracle_mod.count_degenerate(xyz, m) == 0 for m in M)
    _check(oracle_mod, xyz, M, W, H, L)


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_every_single_view_kernel_variant_is_bit_exact(oracle_mod, mode):
    """B == 1 with nested levels can run the staged kernel (0) or the lean kernel (1: RED only, 2: early-z + RED,
    3: shared-memory filter + RED); all must produce the oracle's packed z-buffer bit for bit."""
    from read_b200 import _lib as L
    lib = L.load()
    try:
        L.check(lib.read_set_option(b"raster_mode", mode))
        for n, W, H, Lv, t in ((100_000, 256, 256, 5, 0), (300_000, 128, 64, 3, 4), (4097, 64, 32, 2, 9)):
            xyz, M = scene_and_cams(n, W, H, [t], depth=40.0)      # heavy overdraw: many points per pixel
            _check(oracle_mod, xyz, M, W, H, Lv)
    finally:
        L.check(lib.read_set_option(b"raster_mode", 2))


def test_shared_reciprocal_division_edge_values(oracle_mod):
    """The lean kernel replaces three IEEE divisions by one reciprocal + residual corrections and culls before dividing.
    Points exactly on the frustum planes (|a| == |w|), one ulp either side of them, huge / tiny w and w == 0 must
    give the oracle's result bit for bit (the oracle divides with the C '/' operator)."""
    rng = np.random.default_rng(5)
    base = rng.uniform(-1, 1, size=(4000, 3)).astype(np.float32)
    w = np.exp(rng.uniform(-60, 60, size=4000)).astype(np.float32)
    w[:50] = 0.0
    w[50:100] = np.float32(1e-42)                  # denormal w -> fallback path
    w[100:150] = np.float32(3e38)
    pts = base.copy()
    # x, y, z on the planes and one ulp off them
    pts[:1000, 0] = np.where(rng.random(1000) < 0.5, 1.0, -1.0)
    pts[1000:1500, 0] = np.nextafter(np.float32(1.0), np.float32(2.0))
    pts[1500:2000, 1] = np.nextafter(np.float32(1.0), np.float32(0.0))
    pts[2000:2500, 2] = -1.0
    pts[2500:3000, 2] = np.nextafter(np.float32(-1.0), np.float32(0.0))
    # matrix = diag(1,1,1,0) with w taken from a 4th "coordinate": emulate per-point w by scaling the point and using
    # M[3] = (0,0,0,1) * s ... simpler: fold w into the points (clip = (x*w, y*w, z*w, w)) with M's last row (0,0,1,0)
    # reading w from z: z' = w, and x' = x*w, y' = y*w, with depth row = (0,0,c,0) so cz = c exactly.
    xyz = np.stack([pts[:, 0] * w, pts[:, 1] * w, w], axis=1).astype(np.float32)
    for c in (np.float32(0.25), np.float32(-0.5), np.float32(1.0)):
        M = np.zeros((1, 4, 4), np.float32)
        M[0, 0, 0] = 1; M[0, 1, 1] = 1; M[0, 2, 2] = c; M[0, 3, 2] = 1
        _check(oracle_mod, xyz, M, 64, 48, 1)
        _check(oracle_mod, xyz, M, 1920, 1088, 1)


def test_c2_one_million_points_512(oracle_mod):
    xyz, M = scene_and_cams(1_000_000, 512, 512, [3], depth=250.0, seed=synth.SEED)
    _check(oracle_mod, xyz, M, 512, 512, 4)


def test_unaligned_points_pointer_and_tail(oracle_mod):
    xyz, M = scene_and_cams(5000, 64, 64, [0])
    d = dev()
    big = torch.from_numpy(xyz).to(d)
    sub = big[1:4998]                                     # 12-byte offset: not 16B aligned -> LDG path
    assert sub.data_ptr() % 16 != 0 and sub.is_contiguous()
    pyr = ops.Pyramid(1, 64, 64, 2, d)
    pyr.clear()
    ops.raster_project(pyr, sub, torch.from_numpy(M).to(d))
    gi, gd = ops.zbuf_resolve(pyr, 0)
    oi, od = oracle_mod.pcpr_forward(xyz[1:4998], M, 64, 64)
    np.testing.assert_array_equal(gi.cpu().numpy(), oi)
    np.testing.assert_array_equal(gd.cpu().numpy(), od)


def test_empty_cloud_and_all_culled(oracle_mod):
    gi, gd, _ = render_gpu(np.zeros((0, 3), np.float32), ID, 8, 8, 2)
    assert not gi[0].any() and not gd[1].any()
    far = np.full((100, 3), 50.0, np.float32)
    gi, gd, _ = render_gpu(far, ID, 8, 8, 1)
    assert not gi[0].any() and not gd[0].any()


def test_nan_and_near_plane_are_culled():
    # documented deviations: w == 0 (NaN after division) and d == 0 exactly never enter the z-buffer
    M = ID.copy()
    M[0, 3] = [0, 0, 0, 0]                                # w = 0 for every point
    gi, gd, _ = render_gpu(np.array([[0, 0, 0], [0.5, 0.5, 0.5]], np.float32), M, 4, 4, 1)
    assert not gi[0].any() and not gd[0].any()
    gi, gd, _ = render_gpu(np.array([[9, 9, 9], [0.0, 0.0, -1.0]], np.float32), ID, 4, 4, 1)
    assert not gi[0].any() and not gd[0].any()


def test_deterministic_across_runs():
    xyz, M = scene_and_cams(200_000, 128, 128, [0], depth=30.0)     # heavy overdraw
    a, ad, _ = render_gpu(xyz, M, 128, 128, 4)
    for _ in range(3):
        b, bd, _ = render_gpu(xyz, M, 128, 128, 4)
        for l in range(4):
            np.testing.assert_array_equal(a[l], b[l])
            np.testing.assert_array_equal(ad[l], bd[l])


def test_pcpr_forward_api_contract(oracle_mod):
    xyz, M = scene_and_cams(20_000, 80, 48, [0, 3])
    idx, dep = pcpr.forward(torch.from_numpy(xyz), torch.from_numpy(M), 80, 48, 512)   # CPU in, CPU out
    assert not idx.is_cuda and idx.dtype == torch.float32 and tuple(idx.shape) == (2, 48, 80)
    oi, od = oracle_mod.pcpr_forward(xyz, M, 80, 48)
    np.testing.assert_array_equal(idx.numpy(), oi)
    np.testing.assert_array_equal(dep.numpy(), od)
    i2, d2 = pcpr.forward(torch.from_numpy(xyz).cuda(), torch.from_numpy(M).cuda(), 80, 48, 256)
    assert torch.equal(i2, idx) and torch.equal(d2, dep)
    with pytest.raises(RuntimeError, match="float"):
        pcpr.forward(torch.from_numpy(xyz).double(), torch.from_numpy(M), 8, 8, 512)
    with pytest.raises(RuntimeError, match="contiguous"):
        pcpr.forward(torch.from_numpy(xyz).t().contiguous().t(), torch.from_numpy(M), 8, 8, 512)
    with pytest.raises(RuntimeError, match="batch_size check"):
        pcpr.forward(torch.from_numpy(xyz), torch.from_numpy(M[0]), 8, 8, 512)


def test_myrender_matches_oracle_render(oracle_mod):
    class DS:
        pass
    W, H, L = 96, 64, 5
    ds0, ds1 = DS(), DS()
    xyz0 = synth.street_scene(20_000, depth=50.0, seed=3)
    xyz1 = synth.street_scene(15_000, depth=50.0, seed=4)
    for i, (ds, x) in enumerate(((ds0, xyz0), (ds1, xyz1))):
        ds.id, ds.tgt_sh = i, np.array([W, H])
        ds.input_format = "uv_1d_p1, uv_1d_p1_ds1, uv_1d_p1_ds2, uv_1d_p1_ds3, uv_1d_p1_ds4"
        ds.scene_data = {'pointcloud': {'xyz': x}}
    r = MyRender([ds0, ds1])
    proj, view = synth.camera_batch(W, H, [0, 5, 9])
    ids = torch.tensor([1, 0, 1])
    data = {'input': {'id': ids}, 'proj_matrix': torch.from_numpy(proj), 'view_matrix': torch.from_numpy(view)}
    out, depth = r.render(data)
    assert out['id'] is ids
    keys = [k.strip() for k in ds0.input_format.split(',')]
    for b, (x, did) in enumerate(((xyz1, 1), (xyz0, 0), (xyz1, 1))):
        _, oi, od = oracle_mod.render_pyramid(x, proj[b:b + 1], view[b:b + 1], W, H, L)
        for l, k in enumerate(keys):
            assert tuple(out[k].shape) == (3, 1) + oi[l].shape[2:] and not out[k].is_cuda
            np.testing.assert_array_equal(out[k][b].numpy(), oi[l][0])
            np.testing.assert_array_equal(depth[k][b].numpy(), od[l][0])


def test_full_size_properties_10m_points():
    """BASELINE size (10M points, 1920x1072, L=4): size-independent properties instead of the slow oracle."""
    n, W, H, L = 10_000_000, 1920, 1072, 4
    xyz = synth.street_scene(n)
    proj, view = synth.camera_batch(W, H, [7])
    M = synth.total_matrix(proj, view)
    gi, gd, pyr = render_gpu(xyz, M, W, H, L)
    d = dev()
    x = torch.from_numpy(xyz).to(d)
    m = torch.from_numpy(M).to(d)
    # (1) derived (2x2-min) levels == direct rasterisation of each level on its own
    for l in range(1, L):
        w, h = pyr.sizes[l]
        i, dd = ops.pcpr_forward_device(x, m, w, h)
        assert torch.equal(i.cpu(), torch.from_numpy(gi[l])) and torch.equal(dd.cpu(), torch.from_numpy(gd[l]))
    # (2) sharding: min over two half-clouds (global ids) == the full render, key for key
    full = pyr.buf.clone()
    half = (n // 2 // 1024) * 1024
    pa, pb = ops.Pyramid(1, W, H, L, d), ops.Pyramid(1, W, H, L, d)
    pa.clear(); pb.clear()
    ops.raster_project(pa, x[:half], m, id_base=0)
    ops.raster_project(pb, x[half:], m, id_base=half)
    assert torch.equal(torch.minimum(pa.buf, pb.buf), full)
    # (3) every winner really projects into its pixel with exactly that depth (re-project winners on the host)
    idx0 = gi[0][0]
    ys, xs = np.nonzero(gd[0][0])
    sel = np.random.default_rng(0).choice(len(ys), 20000, replace=False)
    ids = idx0[ys[sel], xs[sel]].astype(np.int64)
    import oracle
    for j in range(0, 20000, 5000):
        k = sel[j:j + 5000]
        pid = idx0[ys[k], xs[k]].astype(np.int64)
        oi, od = oracle.pcpr_forward(xyz[pid], M, W, H)
        # each re-projected winner must hit its own pixel; its depth there can only be <= (another winner may share it)
        assert (od[0][ys[k], xs[k]] == gd[0][0][ys[k], xs[k]]).all()
    # (4) idempotence: re-rendering into the same pyramid changes nothing
    ops.raster_project(pyr, x, m)
    assert torch.equal(pyr.buf, full)
    cov = float((gd[0] != 0).mean())
    assert 0.3 < cov <= 1.0, cov


@pytest.mark.parametrize("dedup", [3, 1, 0, 2])       # 3 = streaming kernel (default); round-1 kernel: 1 = in-warp per-pixel reduction, 2 = neighbour filter, 0 = neither
@pytest.mark.parametrize("n,W,H,L,t,depth,cell", [(100_000, 256, 256, 4, 0, 40.0, 0.25), (300_000, 128, 64, 3, 4, 40.0, 0.5),
                                                    (4097, 64, 32, 2, 9, 60.0, 1.0), (1_000_000, 512, 512, 4, 3, 250.0, 0.25)])
def test_sorted_store_is_bit_exact(oracle_mod, dedup, n, W, H, L, t, depth, cell):
    """The spatially sorted store (points permuted, original ids carried) with / without the in-warp per-pixel reduction
    produces the oracle's index and depth maps bit for bit: the z-buffer is order independent."""
    from read_b200 import _lib as Lb
    lib = Lb.load()
    xyz, M = scene_and_cams(n, W, H, [t], depth=depth)
    d = dev()
    store = ops.SortedPoints(torch.from_numpy(xyz).to(d), cell=cell)
    assert torch.equal(torch.sort(store.perm).values, torch.arange(n, device=d))           # a permutation
    assert torch.equal(store.pts4[:, :3], torch.from_numpy(xyz).to(d)[store.perm])
    try:
        Lb.check(lib.read_set_option(b"raster_stream", 1 if dedup == 3 else 0))
        Lb.check(lib.read_set_option(b"raster_dedup", 1 if dedup == 1 else 0))
        Lb.check(lib.read_set_option(b"raster_nbr_filter", 1 if dedup == 2 else 0))
        pyr = ops.Pyramid(1, W, H, L, d)
        pyr.clear()
        ops.raster_project_sorted(pyr, store, torch.from_numpy(M).to(d))
        ops.raster_derive(pyr)
    finally:
        Lb.check(lib.read_set_option(b"raster_stream", 1))
        Lb.check(lib.read_set_option(b"raster_dedup", 0))
        Lb.check(lib.read_set_option(b"raster_nbr_filter", 0))
    for l, (w, h) in enumerate(oracle_mod.level_sizes(W, H, L)):
        gi, gd = ops.zbuf_resolve(pyr, l)
        oi, od = oracle_mod.pcpr_forward(xyz, M, w, h)
        np.testing.assert_array_equal(gi.cpu().numpy(), oi, err_msg=f"index level {l}")
        np.testing.assert_array_equal(gd.cpu().numpy().view(np.uint32), od.view(np.uint32), err_msg=f"depth level {l}")


@pytest.mark.parametrize("n,B", [(200_000, 3), (5000, 8), (1_000_001, 2)])
def test_sorted_store_multi_view_single_pass(oracle_mod, n, B):
    """read_raster_project_sorted_views: B views rasterised in ONE pass over the store (the multi-GPU frame path) equal the
    oracle view by view; n not a multiple of the 1024-point chunk exercises the partial last bulk copy."""
    W, H, L = 256, 128, 4
    xyz, M = scene_and_cams(n, W, H, list(range(2, 2 + B)), depth=80.0)
    d = dev()
    store = ops.SortedPoints(torch.from_numpy(xyz).to(d))
    pyr = ops.Pyramid(B, W, H, L, d)
    pyr.clear()
    ops.raster_project_sorted(pyr, store, torch.from_numpy(M).to(d))
    ops.raster_derive(pyr)
    for l, (w, h) in enumerate(oracle_mod.level_sizes(W, H, L)):
        gi, gd = ops.zbuf_resolve(pyr, l)
        oi, od = oracle_mod.pcpr_forward(xyz, M, w, h)
        np.testing.assert_array_equal(gi.cpu().numpy(), oi, err_msg=f"index level {l}")
        np.testing.assert_array_equal(gd.cpu().numpy().view(np.uint32), od.view(np.uint32), err_msg=f"depth level {l}")


def test_sorted_store_depth_ties_take_the_lowest_original_id(oracle_mod):
    """Many points with IDENTICAL coordinates (same pixel, same depth bits): lowest original id must win although the sorted
    store places them in one warp (the in-warp reduction resolves ties with a second atomic on the id)."""
    rng = np.random.default_rng(1)
    base = rng.uniform(-0.8, 0.8, size=(200, 3)).astype(np.float32)
    base[:, 2] = rng.uniform(-0.9, 0.9, size=200).astype(np.float32)
    xyz = np.repeat(base, 40, axis=0)                      # 40 copies of each point
    xyz = xyz[rng.permutation(len(xyz))]
    d = dev()
    store = ops.SortedPoints(torch.from_numpy(xyz).to(d), cell=0.05)
    pyr = ops.Pyramid(1, 64, 64, 1, d)
    pyr.clear()
    ops.raster_project_sorted(pyr, store, torch.from_numpy(ID).to(d))
    gi, gd = ops.zbuf_resolve(pyr, 0)
    oi, od = oracle_mod.pcpr_forward(xyz, ID, 64, 64)
    np.testing.assert_array_equal(gi.cpu().numpy(), oi)
    np.testing.assert_array_equal(gd.cpu().numpy(), od)


def test_sorted_store_full_size_equals_unsorted_render():
    n, W, H, L = 10_000_000, 1920, 1088, 4
    xyz = synth.street_scene(n)
    proj, view = synth.camera_batch(W, H, [7])
    d = dev()
    m = torch.from_numpy(synth.total_matrix(proj, view)).to(d)
    x = torch.from_numpy(xyz).to(d)
    a, b = ops.Pyramid(1, W, H, L, d), ops.Pyramid(1, W, H, L, d)
    a.clear(); b.clear()
    ops.raster_project(a, x, m)
    store = ops.SortedPoints(x)
    ops.raster_project_sorted(b, store, m)
    ops.raster_derive(b)
    assert torch.equal(a.buf, b.buf)
