"""world_size-2 gloo test of the multi-GPU host logic on CPU: sharding + min-reduce of packed keys gives the
same z-buffer as the single-process sequential oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from read_b200 import dist as rdist, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(index, depth):
    """(depth|id) int64 keys from an oracle render of ONE shard whose ids are already global."""
    d = depth.view(np.uint32).astype(np.int64) << 32
    k = d | index.astype(np.int64)
    k[depth == 0] = rdist.EMPTY_KEY
    return k


def _worker(rank, world, port, xyz, M, w, h, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        start, count = rdist.shard_range(xyz.shape[0], rank, world, align=64)
        # render the shard with the oracle; ids inside the shard are local -> shift to global where non-empty
        idx, dep = oracle.pcpr_forward(xyz[start:start + count], M, w, h)
        gidx = np.where(dep != 0, idx + start, 0).astype(np.float32)
        keys = torch.from_numpy(_pack(gidx, dep))
        # the frame path's collective: view r's plane goes to rank r only (B == world views)
        mine = torch.empty(keys[0].numel(), dtype=torch.int64)
        rdist.reduce_scatter_min_(mine, keys.reshape(-1).clone())
        ret[("rs", rank)] = mine.numpy().copy()
        rdist.allreduce_min_(keys)
        ret[rank] = keys.numpy().copy()
    finally:
        dist.destroy_process_group()


def test_sharded_min_reduce_equals_sequential(oracle_mod):
    xyz = synth.street_scene(5000, depth=40.0, seed=5)
    proj, view = synth.camera_batch(48, 32, [0, 4])
    M = synth.total_matrix(proj, view)
    idx, dep = oracle_mod.pcpr_forward(xyz, M, 48, 32)
    want = _pack(idx, dep)
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, xyz, M, 48, 32, ret), nprocs=world, join=True)
    for r in range(world):
        np.testing.assert_array_equal(ret[r], want)
        np.testing.assert_array_equal(ret[("rs", r)], want[r].reshape(-1))       # reduce-scatter: rank r holds view r
