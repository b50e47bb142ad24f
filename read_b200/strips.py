"""Halo arithmetic of the strip-parallel refinement net (READ/models/unet.py:202-285 is the dataflow).

The frame is cut into ``world`` horizontal strips.  Rank r holds, of every activation tensor at level l (scale 2^-l), the crop
``[y0/2^l - h_l, y1/2^l + h_l)`` with ``h_l = HALO0 >> l`` halo rows per side (none at the image border), aligned so that every
resampling of the net (stride-2 convs, nearest x2^k up / down, bilinear x4) maps crop rows onto crop rows exactly as it maps
image rows onto image rows: the existing kernels run on a crop as on a smaller image.  What differs is the outermost rows: a
kernel pads with zeros where the neighbouring strip's data should be, so every spatial op makes some halo rows wrong.
``validity`` counts the halo rows (per side) that still hold true values; the rules below are the exact per-op bookkeeping
(taking the smaller of the top / bottom bound so that all ranks take identical decisions), and the engine exchanges a tensor's
halo with the neighbours (restoring ``h_l``) right before an op that would otherwise read an invalid row.

Pure functions, no torch / CUDA: tests/test_strips.py replays the plan with torch CPU convolutions on two emulated ranks and
checks the stitched result against the full-frame computation.
"""

HALO0 = 16          # halo rows at full resolution: one 16-row tile; 2 rows at 1/8 resolution


def halo_rows(level):
    return HALO0 >> level


def src_validity(v, mode, factor, h_in):
    """Valid halo rows a source contributes AT THE CONV'S INPUT RESOLUTION.  ``v``: the source tensor's own valid rows; ``mode``
    'id' | 'down' (nearest, ::factor) | 'up' (nearest, x factor) | 'bil4' (bilinear x4, align_corners=False); ``h_in``: halo
    rows of the conv-input level (cap)."""
    if mode == "id":
        return v
    if mode == "down":            # input row j samples source row factor * j: halo row j is valid iff factor * j <= v
        return v // factor
    if mode == "up":              # input rows (c - 1) * factor + 1 .. c * factor copy the source's c-th halo row
        return min(v * factor, h_in)
    if mode == "bil4":            # input halo row j blends source rows up to ceil((j + 2) / 4) deep
        return max(0, min(4 * v - 2, h_in))
    raise ValueError(mode)


def conv_need(k):
    """Halo rows of its (resampled) input a conv reads for its INTERIOR output rows."""
    return 0 if k == 1 else 1


def conv_out_validity(v_in, k, stride):
    """Valid halo rows of the output given ``v_in`` valid rows of the input (at input resolution)."""
    if k == 1:
        return v_in if stride == 1 else v_in // 2
    if stride == 1:               # 3x3, pad 1: one row lost per side
        return max(v_in - 1, 0)
    # 3x3 / 4x4, stride 2, pad 1: output halo row j reads input rows down to 2j + 1 beyond the edge
    return max((v_in - 1) // 2, 0)


def strip_rows(H, world, rank, level=0):
    """(first local row's global index, local row count, top halo, bottom halo) of rank's crop at ``level``."""
    assert H % (HALO0 * world) == 0, f"frame height {H} must be a multiple of {HALO0 * world} for {world} strips"
    S = H // world
    h = halo_rows(level)
    top = h if rank > 0 else 0
    bot = h if rank < world - 1 else 0
    y0 = (rank * S) >> level
    return y0 - top, (S >> level) + top + bot, top, bot
