"""The built library really contains the Blackwell instructions DESIGN.md claims (checked on the SASS of libread_b200.so, no GPU needed):
tcgen05.mma (UTCHMMA, and its cta_group::2 form in the CTA-pair kernels), TMEM loads (LDTM), TMA tile loads / stores (UTMALDG / UTMASTG),
the multicast commit of the pair kernels (UTCBAR.2CTA.MULTICAST) and the bulk copy of the streaming rasterizer (UBLKCP); and that the
work-skipping diagnostic knobs are absent from the shipped build.  Mnemonics as listed in /opt/skills/guides/B200_PROFILING.md."""
import collections
import os
import re
import shutil
import subprocess

import pytest

from read_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libread_b200.so not built")
    out = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    per_fn, cur = collections.defaultdict(collections.Counter), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            per_fn[cur][m.group(1)] += 1
    return per_fn


def _kernels(per_fn, needle):
    ks = {k: v for k, v in per_fn.items() if needle in k}
    assert ks, f"no kernel named *{needle}* in the library"
    return ks


def _has(counter, prefix):
    return sum(n for op, n in counter.items() if op.startswith(prefix))


def test_single_cta_conv_kernels_use_tcgen05_tmem_and_tma(sass):
    for name, ops in _kernels(sass, "gated_conv_tc_kernel").items():
        assert _has(ops, "UTCHMMA") > 0 and _has(ops, "LDTM") > 0 and _has(ops, "UTMALDG") > 0, name
        assert _has(ops, "UTCHMMA.2CTA") == 0, name
    lean = {k: v for k, v in _kernels(sass, "gated_conv_tc_kernel").items() if "ELi640E" in k}
    assert lean and all(_has(v, "UTMASTG") > 0 for v in lean.values()), "lean epilogues stage their items for TMA stores"


def test_cta_pair_kernels_use_cta_group_2(sass):
    for needle in ("gated_conv_tc2_kernel", "gated_conv_tc2s_kernel"):
        for name, ops in _kernels(sass, needle).items():
            assert _has(ops, "UTCHMMA.2CTA") > 0, name                       # tcgen05.mma.cta_group::2
            assert _has(ops, "UTCHMMA") == _has(ops, "UTCHMMA.2CTA"), name    # ... and nothing else
            assert _has(ops, "UTMALDG.4D.2CTA") > 0 and _has(ops, "UTMALDG.2D.2CTA") > 0, name
            assert _has(ops, "UTCBAR.2CTA.MULTICAST") > 0, name               # one commit releases both CTAs
            assert _has(ops, "LDTM") > 0, name
    assert all(_has(v, "UTMASTG") > 0 for v in _kernels(sass, "gated_conv_tc2_kernel").values())


def test_gather_kernel_and_rasterizer(sass):
    for name, ops in _kernels(sass, "gated_conv_tc_gather_kernel").items():
        assert _has(ops, "UTCHMMA") > 0 and _has(ops, "LDTM") > 0 and _has(ops, "LDGSTS") > 0, name     # cp.async-gathered A operand
    for name, ops in _kernels(sass, "raster_stream_kernel").items():
        assert _has(ops, "UBLKCP") > 0 and _has(ops, "SYNCS") > 0, name        # cp.async.bulk ring + mbarriers
        assert _has(ops, "RED") + _has(ops, "ATOM") > 0, name                 # 64-bit min into the packed z-buffer


def test_shipped_library_rejects_the_diagnostic_knobs():
    lib = _lib.load()
    for knob in (b"tc_debug", b"tcg_debug"):
        assert lib.read_set_option(knob, 1) != 0                              # unknown to the shipped build (READ_DIAG only)
    assert lib.read_set_option(b"raster_mode", 4) != 0 and lib.read_set_option(b"raster_mode", 5) != 0
    assert lib.read_set_option(b"raster_mode", 2) == 0 and lib.read_set_option(b"tc_pdl", 1) == 0
    assert lib.read_set_option(b"no_such_option", 1) != 0
