"""The C-ABI library loads and exports every symbol include/read_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from read_b200 import _lib


def _declared():
    src = open(os.path.join(ROOT, "include", "read_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(read_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_is_built_and_loads():
    lib = _lib.load()
    assert lib.read_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/read_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared, set(_lib.EXPORTS) ^ set(declared)


def test_pyramid_geometry_host_functions():
    lib = _lib.load()
    # myrender.py:33-34 sizes; kitti6 viewport and 1080p
    w, h = ctypes.c_int(), ctypes.c_int()
    lib.read_level_size(1216, 368, 4, ctypes.byref(w), ctypes.byref(h))
    assert (w.value, h.value) == (76, 23)
    lib.read_level_size(1920, 1080, 4, ctypes.byref(w), ctypes.byref(h))
    assert (w.value, h.value) == (120, 67)
    P = sum((1920 >> l) * (1072 >> l) for l in range(4))
    assert P == 2733600                                  # P for C3
    assert lib.read_pyramid_entries(1, 1920, 1072, 4) == P
    assert lib.read_pyramid_entries(3, 1920, 1072, 4) == 3 * P
    assert lib.read_pyramid_level_offset(2, 64, 32, 2) == 2 * (64 * 32 + 32 * 16)
    assert lib.read_pyramid_entries(1, 64, 64, 99) == -1


def test_direct_mask_nested_vs_odd_levels():
    lib = _lib.load()
    assert lib.read_raster_direct_mask(1920, 1072, 4) == 0b0001       # every level an exact halving
    assert lib.read_raster_direct_mask(1920, 1080, 5) == 0b10001      # 135 -> 67 is not
    assert lib.read_raster_direct_mask(1216, 368, 5) == 0b00001
    assert lib.read_raster_direct_mask(100, 50, 3) == 0b101           # 50,25,12 / 100,50,25 : level 2 odd


def test_generic_packing_geometry():
    lib = _lib.load()
    assert lib.read_generic_npad(3) == 64 and lib.read_generic_npad(32) == 64 and lib.read_generic_npad(56) == 128
    assert lib.read_tc_weight_elems(32, 32, 3) == 9 * 32 * 64
    assert lib.read_tc_weight_elems(3, 32, 3) == 9 * 32 * 16          # final layer: Cout padded to 8 (N = 16)
    assert lib.read_tc_weight_elems(56, 32, 1) == -1                  # not a TMA-kernel shape (gather kernel takes it)
    assert lib.read_tcg_weight_elems(56, 32, 1) == 64 * 112           # K padded to one 64-block, N = 2*56


def test_conv_validation_errors_without_gpu():
    lib = _lib.load()
    d = _lib.ReadConvDesc()
    plan = _lib.c_vp()
    rc = lib.read_conv_plan_create(ctypes.byref(d), ctypes.byref(plan))
    assert rc == -1 and b"conv" in lib.read_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc)
