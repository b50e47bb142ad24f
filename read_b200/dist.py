"""Multi-GPU rendering: point-sharded rasterization joined by ONE min-reduction of the packed z-buffer.

The z-test is an associative, commutative min over packed (depth|id) keys, so each rank rasterises its slice of
the cloud (keeping GLOBAL point ids via ``id_base``) into a full-size pyramid and a single collective yields the
same winners on every rank, bit for bit.  Only the levels that were rasterised with direct
atomics are exchanged (level 0 alone when the pyramid nests: 16.5 MB of the 21.9 MB at 1920x1072); coarser
levels are re-derived locally after the reduce.

One process per GPU (torch.distributed, backend "nccl"; "gloo" on CPU for the host-logic tests).  The
refinement net does not shard (receptive field spans the frame): frames are refined frame-parallel, rank r
refines view r of each step's batch of ``world_size`` views.
"""
import torch
import torch.distributed as dist

EMPTY_KEY = 0x7FFFFFFFFFFFFFFF


def shard_range(n_points, rank, world_size, align=1024):
    """Contiguous slice [start, start+count) of the id range owned by ``rank``; boundaries are multiples of
    ``align`` points so every shard's xyz pointer stays 16-byte aligned for the bulk-TMA loader."""
    blocks = (n_points + align - 1) // align
    per, rem = divmod(blocks, world_size)
    b0 = rank * per + min(rank, rem)
    b1 = b0 + per + (1 if rank < rem else 0)
    start = min(b0 * align, n_points)
    stop = min(b1 * align, n_points)
    return start, stop - start


def reduce_span(offsets, sizes, B, direct_levels):
    """[lo, hi) entry range of the pyramid buffer covering all directly rasterised levels."""
    lo = min(offsets[l] for l in direct_levels)
    hi = max(offsets[l] + B * sizes[l][0] * sizes[l][1] for l in direct_levels)
    return lo, hi


def allreduce_min_(keys, group=None):
    """In-place elementwise min of packed int64 keys across ranks (ncclMin / gloo MIN)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return keys


def reduce_scatter_min_(out, keys, group=None):
    """Elementwise min across ranks of ``keys`` ([world * n] packed int64 keys: view r's level-0 plane at [r*n, (r+1)*n)), rank r
    receiving only ITS view's plane in ``out`` [n]: half the bytes of an all-reduce and 1/world of its output traffic
    (every rank consumes one view).  NCCL: ncclReduceScatter(int64, min); gloo has no reduce-scatter and
    falls back to all-reduce + slice (host-logic tests only)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out.copy_(keys[:out.numel()])
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    assert keys.numel() == world * out.numel(), (keys.numel(), world, out.numel())
    if dist.get_backend(group) == "gloo":
        tmp = keys.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.MIN, group=group)
        out.copy_(tmp[rank * out.numel():(rank + 1) * out.numel()])
    else:
        dist.reduce_scatter_tensor(out, keys, op=dist.ReduceOp.MIN, group=group)
    return out


def render_sharded(pyr, xyz_shard, id_base, total_m, group=None):
    """Project this rank's shard into ``pyr`` (cleared here), min-reduce, derive nested levels.
    After the call every rank holds the identical, complete pyramid.  ``xyz_shard``: this rank's slice of the [N,3] cloud
    (global ids via ``id_base``) or a ``ops.SortedPoints`` shard (a spatial tile; ids travel with the points)."""
    from . import ops
    pyr.clear()
    if isinstance(xyz_shard, ops.SortedPoints):
        ops.raster_project_sorted(pyr, xyz_shard, total_m)
    else:
        ops.raster_project(pyr, xyz_shard, total_m, id_base=id_base, derive=False)
    lo, hi = reduce_span(pyr.offsets, pyr.sizes, pyr.B, pyr.direct_levels())
    allreduce_min_(pyr.buf[lo:hi], group)
    ops.raster_derive(pyr)
    return pyr


class ShardedFrameStream:
    """Throughput mode: ``world`` views per step, rank r refines view r, with the rasterizer ONE STEP AHEAD of
    the net.  A step is: this rank's spatial tile rasterised for all ``world`` views in one pass, ONE reduce-scatter(min) of the
    level-0 planes (rank r receives view r), fused resolve + gather into the net's inputs, the net.  The first three touch only the
    pyramid and ``recv``; the net touches only its own buffers - so while the net of step i runs on the caller's stream, a side
    stream already clears, rasterises and reduce-scatters step i+1 (``m_next``).  Per step the critical path is
    max(gather + net, raster + collective) instead of their sum.

    ``step(m_dev, m_next)``: m_dev / m_next are ``[world,4,4]`` device tensors (proj @ inv(view) per view); pass as ``m_next`` the
    very tensor the next call will pass as ``m_dev`` (or None).  Returns the engine's output ``[1,3,H,W]`` for this rank's view."""

    def __init__(self, store, tex_nd, engine, W, H, n_levels, layout, group=None):
        from . import ops
        self.ops, self.store, self.tex_nd, self.eng, self.layout, self.group = ops, store, tex_nd, engine, layout, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        dev = tex_nd.device
        self.pyr = ops.Pyramid(self.world, W, H, n_levels, dev)
        self.plane = W * H
        self.recv = torch.empty(self.plane, dtype=torch.int64, device=dev)
        self.side = torch.cuda.Stream(device=dev)
        self.rs_done, self.gather_done = torch.cuda.Event(), torch.cuda.Event()
        self._pending = None                        # the matrices whose reduced plane is (being) produced into recv

    def _raster_and_reduce(self, m_dev):
        """On the CURRENT stream: clear level 0 of all views, one pass over the shard, reduce-scatter(min) into recv."""
        from . import _lib as L
        L.check(L.load().read_zbuf_clear(self.pyr.buf.data_ptr(), self.world * self.plane, L.stream_ptr()))
        self.ops.raster_project_sorted(self.pyr, self.store, m_dev)
        reduce_scatter_min_(self.recv, self.pyr.buf[:self.world * self.plane], self.group)

    def step(self, m_dev, m_next=None):
        cur = torch.cuda.current_stream()
        if self._pending is None or self._pending.data_ptr() != m_dev.data_ptr():
            start = torch.cuda.Event()
            start.record(cur)
            with torch.cuda.stream(self.side):      # no look-ahead for this step: do it now, still on the side stream (stream order
                self.side.wait_event(start)         # after whatever look-ahead was in flight)
                self._raster_and_reduce(m_dev)
                self.rs_done.record(self.side)
        cur.wait_event(self.rs_done)
        r = self.rank
        self.pyr.buf[r * self.plane:(r + 1) * self.plane].copy_(self.recv)
        self.ops.pyramid_resolve_gather(self.tex_nd, self.pyr, self.eng.inputs, self.layout, view0=r, nviews=1)
        self.gather_done.record(cur)
        self._pending = None
        if m_next is not None:
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.gather_done)       # the pyramid and recv are free again
                self._raster_and_reduce(m_next)
                self.rs_done.record(self.side)
            m_next.record_stream(self.side)
            self._pending = m_next
        return self.eng.run()


class StripFrameRenderer:
    """Latency mode: ``world`` GPUs cooperate on ONE frame.  Rank r rasterises its spatial tile of the
    scene, ONE all-reduce(min) makes the packed level-0 z-buffer complete on every rank, every rank gathers the feature pyramid
    and refines its horizontal strip with ``engine.StripEngine`` (halo rows exchanged layer by layer over NVLink peer memory);
    the strips are all-gathered into the full frame on every rank."""

    def __init__(self, store_shard, texture_point_major, state_dict, W, H, device, group=None):
        from . import ops, _lib as L
        from .engine import StripEngine
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.group, self.W, self.H, self.device = group, W, H, device
        self.store, self.tex = store_shard, texture_point_major
        self.pyr = ops.Pyramid(1, W, H, 4, device)
        self.eng = StripEngine(state_dict, H, W, device, self.rank, self.world, group=group)
        self.full_feats = [torch.empty((1, H >> l, W >> l, 8), dtype=torch.bfloat16, device=device) for l in range(4)]
        self.strips_out = torch.empty((self.world, 3, H // self.world, W), dtype=torch.float32, device=device)
        self._ops, self._L = ops, L

    def render(self, total_m):
        """total_m [1,4,4] cuda f32 (the same on every rank) -> [3,H,W] f32 frame, identical on every rank."""
        ops, L = self._ops, self._L
        pyr, plane = self.pyr, self.W * self.H
        L.check(L.load().read_zbuf_clear(pyr.buf.data_ptr(), plane, L.stream_ptr()))
        ops.raster_project_sorted(pyr, self.store, total_m)
        allreduce_min_(pyr.buf[:plane], self.group)
        ops.pyramid_resolve_gather(self.tex, pyr, self.full_feats, L.FEAT_NHWC_BF16)
        self.eng.set_inputs_from_full(self.full_feats)
        self.eng.run()
        dist.all_gather_into_tensor(self.strips_out, self.eng.output_interior[0].contiguous(), group=self.group)
        return self.strips_out.permute(1, 0, 2, 3).reshape(3, self.H, self.W)
