"""Known-answer tests pinning the sequential z-buffer oracle (oracle/zbuffer.c) to hand-derived values
(SURVEY.md §8c(1)): the reference ships no golden vectors, so these are derived from the arithmetic of
point_render.cu:107-167 and READ/gl/utils.py:123-150 by hand."""
import numpy as np
import pytest

from read_b200 import synth

ID = np.eye(4, dtype=np.float32)[None]


def test_identity_matrix_center_pixel(oracle_mod):
    # M = I: camp = (x,y,z); u = W(x+1)/2, v = H(1-y)/2, d = (z+1)/2
    xyz = np.array([[0.0, 0.0, 0.0]], np.float32)
    idx, dep = oracle_mod.pcpr_forward(np.concatenate([np.zeros((1, 3), np.float32) + 5, xyz]), ID, 8, 4)
    # point 0 is outside (x=5 > 1), point 1 lands at (u,v) = (4,2), depth 0.5
    assert idx.shape == (1, 4, 8)
    want = np.zeros((4, 8), np.float32)
    want[2, 4] = 1.0
    np.testing.assert_array_equal(idx[0], want)
    assert dep[0, 2, 4] == np.float32(0.5) and dep.sum() == np.float32(0.5)


def test_depth_test_and_tie_break(oracle_mod):
    # three points on the same pixel: farther, nearer, and an exact tie with the nearer one (higher id loses)
    xyz = np.array([[9, 9, 9], [0.1, 0.1, 0.5], [0.1, 0.1, -0.25], [0.1, 0.1, -0.25], [0.1, 0.1, 0.0]], np.float32)
    idx, dep = oracle_mod.pcpr_forward(xyz, ID, 10, 10)
    x, y = int(10 * 1.1 * 0.5), int(10 * 0.9 * 0.5)
    assert idx[0, y, x] == 2.0
    assert dep[0, y, x] == np.float32((np.float32(-0.25) + np.float32(1)) * np.float32(0.5))
    assert (idx != 0).sum() == 1


def test_frustum_planes_inclusive_and_pixel_reject(oracle_mod):
    # |coord| == 1 passes the cull (strict comparisons, :139); x=+1 -> u=W -> rejected by xx>=W (:147);
    # x=-1 -> pixel 0; y=+1 -> v=0 -> row 0; y=-1 -> v=H -> rejected.
    xyz = np.array([[7, 7, 7], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1]], np.float32)
    idx, dep = oracle_mod.pcpr_forward(xyz, ID, 4, 4)
    assert idx[0, 2, 0] == 2.0          # x = -1
    assert idx[0, 0, 2] == 3.0          # y = +1
    assert idx[0, 2, 2] == 5.0 and dep[0, 2, 2] == 1.0   # z = +1 -> d = 1, only point on that pixel
    assert (idx != 0).sum() == 3


def test_empty_is_index0_depth0_and_point0_indistinguishable(oracle_mod):
    # "0 denote empty" (:158): a pixel won by point 0 has index 0 but non-zero depth
    xyz = np.array([[0, 0, 0]], np.float32)
    idx, dep = oracle_mod.pcpr_forward(xyz, ID, 2, 2)
    assert idx.sum() == 0 and dep[0, 1, 1] == 0.5 and (dep != 0).sum() == 1


def test_empty_cloud_and_batch(oracle_mod):
    idx, dep = oracle_mod.pcpr_forward(np.zeros((0, 3), np.float32), np.repeat(ID, 3, 0), 5, 3)
    assert idx.shape == (3, 3, 5) and not idx.any() and not dep.any()


def test_projection_matches_hand_computation(oracle_mod):
    # pinhole camera at origin looking down -z (GL), K from synth.intrinsics: a point on the optical axis
    # lands at (cx, cy); depth follows get_proj_matrix (READ/gl/utils.py:123-150) with near .1, far 1000.
    W, H = 64, 32
    proj, view = synth.camera_batch(W, H, [0])
    M = synth.total_matrix(proj, view)
    xyz = np.array([[50, 50, 50], [0.0, 0.0, -10.0]], np.float32)      # point 0 is behind the camera
    idx, dep = oracle_mod.pcpr_forward(xyz, M, W, H)
    assert idx[0, H // 2, W // 2] == 1.0 and (idx != 0).sum() == 1
    n, f, z = 0.1, 1000.0, 10.0
    z_ndc = ((f + n) / (f - n) * z - 2 * f * n / (f - n)) / z
    assert abs(float(dep[0, H // 2, W // 2]) - (z_ndc + 1) / 2) < 1e-6


def test_behind_camera_is_culled(oracle_mod):
    proj, view = synth.camera_batch(32, 32, [0])
    M = synth.total_matrix(proj, view)
    xyz = np.array([[0, 0, 5.0], [0.3, -0.2, 50.0], [0, 0, -0.05]], np.float32)   # behind, behind, nearer than znear
    idx, dep = oracle_mod.pcpr_forward(xyz, M, 32, 32)
    assert not idx.any() and not dep.any()


def test_level_sizes_match_reference_rule(oracle_mod):
    # src/READ/gl/myrender.py:33-34 with the kitti6 viewport (downloads/kitti6.yaml:1) and 1080p
    assert oracle_mod.level_sizes(1216, 368, 5) == [(1216, 368), (608, 184), (304, 92), (152, 46), (76, 23)]
    assert oracle_mod.level_sizes(1920, 1080, 5)[3:] == [(240, 135), (120, 67)]


def test_odd_level_is_not_a_2x2_reduction(oracle_mod):
    # 135 -> 67 rows is not an exact halving: a point in fine row 134 maps to coarse row 66, not 67.
    W, H = 8, 135
    xyz = np.array([[9, 9, 9], [0.0, -0.999, 0.0]], np.float32)
    i0, _ = oracle_mod.pcpr_forward(xyz, ID, W, H)
    i1, _ = oracle_mod.pcpr_forward(xyz, ID, W // 2, 67)
    y0 = int(np.nonzero(i0[0])[0][0])
    y1 = int(np.nonzero(i1[0])[0][0])
    assert y0 == 134 and y1 == 66
