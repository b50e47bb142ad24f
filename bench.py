#!/usr/bin/env python
"""Benchmark of the READ per-frame render hot path on B200 (driver contract: one JSON line on stdout).

    python bench.py --gpus 1 --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus 1 ...            # the reference's CPU path (oracle port) on host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame of the 10 M-point street scene at 1920x1080 (rendered at 1920x1088 = padded to the %16 the
net needs, READ/gl/nn.py:107-109): clear + project/cull/z-resolve all points into the 4-level packed pyramid,
gather the descriptor feature pyramid, run the full 99-layer gated-conv refinement net -> RGB frame.
With N > 1 GPUs the cloud is sharded by point range, every step renders N camera views together (one pass over
each shard, ONE NCCL min-reduce of the packed level-0 z-buffers) and rank r refines view r (frame-parallel net):
N frames per step, weak scaling.

Outputs (see DESIGN.md "Measurement"): value = frames/s with all inputs resident in HBM; e2e = frames/s through
the public call with the camera matrices coming from pinned HOST memory and the RGB frame copied back to pinned
HOST memory inside the timed region; roofline = dominant kernel (tcgen05 gated conv, tensor bound) measured live
with CUDA events; roofline_raster = the rasterizer against HBM bandwidth; cpu_baseline = the oracle port on the
host cores (bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 10_000_000
W, H, LEVELS = 1920, 1088, 4
H_NAMED = 1080
METRIC = "frames/sec @1920x1080, 10M pts"
WORKLOAD = ("synthetic 10M-point street scene, 1920x1080 (rendered 1920x1088: padded to %16, crop), "
            "L=4 pyramid, descriptor dim 8, full MIMO-UNet refine")


def measured_traffic():
    """DRAM bytes per launch from the committed ncu capture (profiles/r01_traffic.json); None if absent."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v == "Active":
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- CPU arm
def host_threads():
    """Threads the CPU arm may use: affinity mask, capped by the cgroup CPU quota when there is one."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


def pick_torch_threads(sd):
    """torch's CPU convolutions do not scale to every core of a large shared host (oversubscription made a 128-thread
    run 25x slower than an 8-thread one): time one mid-size gated conv at a few thread counts and keep the fastest,
    so the CPU baseline is the best the host can do rather than an artefact."""
    import torch
    from oracle import unet_ref
    avail = host_threads()
    cands = sorted({c for c in (avail, 64, 32, 16, 8) if c <= avail})
    x = torch.rand(1, 64, 256, 512)
    best, best_t = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            unet_ref.basic_conv(sd, "Encoder.1.layers.0.main.0", x, 3)
            t0 = time.perf_counter()
            for _ in range(3):
                unet_ref.basic_conv(sd, "Encoder.1.layers.0.main.0", x, 3)
            t = time.perf_counter() - t0
        if best_t is None or t < best_t:
            best, best_t = c, t
    torch.set_num_threads(best)
    return best, avail


def cpu_reference_frame(xyz, tex_cn, sd, threads):
    """One bounded sample of the reference's CPU path (oracle port): full-size rasterisation of all 4 levels
    (sequential z-buffer, one host thread per level) + gather and the refinement net on a 512x256 window of
    the feature pyramid (torch CPU, all host threads), extrapolated by pixel count to the full frame."""
    import torch
    import oracle
    from oracle import unet_ref
    from read_b200 import synth
    proj, view = synth.camera_batch(W, H, [7])
    t0 = time.perf_counter()
    _, idx, _ = oracle.render_pyramid(xyz, proj, view, W, H, LEVELS, threads=LEVELS)
    t_raster = time.perf_counter() - t0
    cw, ch = 1024, 512
    x0, y0 = (W - cw) // 2, (H - ch) // 2
    t0 = time.perf_counter()
    with torch.no_grad():
        feats = []
        for l in range(LEVELS):
            m = torch.from_numpy(idx[l][:, :, y0 >> l:(y0 + ch) >> l, x0 >> l:(x0 + cw) >> l].copy())
            feats.append(unet_ref.point_texture(tex_cn, m))
        out = unet_ref.unet_forward(sd, feats)
    t_net_crop = time.perf_counter() - t0
    scale = (W * H) / float(cw * ch)
    t_frame = t_raster + t_net_crop * scale
    return t_frame, {"raster_s": t_raster, "net_crop_s": t_net_crop, "crop": [cw, ch], "scale": scale,
                     "out_mean_abs": float(out.abs().mean())}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import oracle
    from read_b200 import synth
    oracle.build()
    sd = synth.synth_state_dict(synth.SEED)
    cores, avail = pick_torch_threads(sd)
    xyz = synth.street_scene(N_POINTS)
    g = torch.Generator().manual_seed(synth.SEED)
    tex = torch.rand((1, 8, N_POINTS), generator=g)
    steps = max(1, min(args.steps, 3))      # each step is ~10-30 s of CPU work: keep the arm within minutes
    warm = 1 if args.warmup > 0 else 0
    for _ in range(warm):
        cpu_reference_frame(xyz, tex, sd, cores)
    ts, info = [], None
    for _ in range(steps):
        t, info = cpu_reference_frame(xyz, tex, sd, cores)
        ts.append(t)
    t_frame = float(np.median(ts))
    fps = 1.0 / t_frame
    sample = (f"per step: full-size sequential z-buffer of all {LEVELS} levels over 10M points (1 thread/level) + "
              f"gather + refinement net on a {info['crop'][0]}x{info['crop'][1]} window x{info['scale']:.2f} "
              f"(pixel-count extrapolation); {steps} step(s), median; torch threads {cores} = fastest of a sweep up to the {avail} available")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": t_frame * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_points": N_POINTS, "width": W, "height": H, "levels": LEVELS},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                             "raster_s": info["raster_s"], "net_crop_s": info["net_crop_s"]},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from read_b200 import synth, ops, _lib as L, dist as rdist
    from read_b200.unet import UNet
    from read_b200.texture import PointTexture
    from read_b200.compose import NetAndTexture

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line (NCCL prints its banner there)
        dist.init_process_group("nccl", device_id=dev)
    L.require_device(local)
    pk = peaks()

    # ---- scene state (loaded once, like MyRender.update_ds / load_textures): resident in HBM
    xyz_np = synth.street_scene(N_POINTS)
    start, count = rdist.shard_range(N_POINTS, rank, world)
    # the spatially sorted store built at scene load (ops.SortedPoints: original ids travel with the points); with N GPUs
    # rank r keeps the r-th contiguous range of the Morton order = a compact spatial tile of the scene
    full_store = ops.SortedPoints(torch.from_numpy(xyz_np).to(dev))
    store = full_store.shard(start, count) if world > 1 else full_store
    if world > 1:
        store.pts4 = store.pts4.clone()       # keep only this rank's tile resident
        store.perm = store.perm.clone()
        del full_store
        torch.cuda.empty_cache()
    g = torch.Generator().manual_seed(synth.SEED)
    tex = PointTexture(8, N_POINTS)
    with torch.no_grad():
        tex.texture_.copy_(torch.rand((1, 8, N_POINTS), generator=g))
    net = UNet()
    net.load_state_dict(synth.synth_state_dict(synth.SEED), strict=True)
    net.precision = args.precision
    model = NetAndTexture(net, {0: tex}, 1)
    model.load_textures(0)
    model.to(dev).eval()
    B = world                                   # views per step
    eng = net.engine(1, H, W, dev)              # each rank refines ONE view per step
    tex_nd = tex.point_major()
    layout = L.FEAT_NHWC_BF16 if eng.bf16 else L.FEAT_NHWC_F32
    pyr = ops.Pyramid(B, W, H, LEVELS, dev)

    n_poses = 64
    total = args.warmup + args.steps
    pose_ts = [[(s * B + v) % n_poses for v in range(B)] for s in range(total)]
    mats_host = torch.empty((total, B, 4, 4), dtype=torch.float32).pin_memory()
    for s in range(total):
        proj, view = synth.camera_batch(W, H, pose_ts[s])
        mats_host[s] = torch.from_numpy(synth.total_matrix(proj, view))
    mats_dev = mats_host.to(dev)
    frame_host = torch.empty((3, H, W), dtype=torch.float32).pin_memory()
    # per-rank view of the batch pyramid for the gather (view `rank` of each level)
    lvl_views = []
    for l in range(LEVELS):
        w_l, h_l = pyr.sizes[l]
        lvl_views.append((pyr.offsets[l] + rank * w_l * h_l, w_l, h_l))

    lib = L.load()

    def step(m_dev):
        """m_dev [B,4,4] on device -> eng.output [1,3,H,W] on device."""
        if world == 1:
            # level 0 is left cleared by the previous frame's fused resolve (reset_level0)
            ops.raster_project_sorted(pyr, store, m_dev)
            ops.pyramid_resolve_gather(tex_nd, pyr, eng.inputs, layout, reset_level0=True)
        else:
            L.check(lib.read_zbuf_clear(pyr.buf.data_ptr(), pyr.B * W * H, L.stream_ptr()))     # level 0 of all views
            ops.raster_project_sorted(pyr, store, m_dev)
            rdist.allreduce_min_(pyr.buf[:pyr.B * W * H])                                        # ONE collective per step
            ops.pyramid_resolve_gather(tex_nd, pyr, eng.inputs, layout, view0=rank, nviews=1)
        return eng.run()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, first):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(first + s)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def resident_step(s):
        step(mats_dev[s])

    def e2e_step(s):
        m = mats_host[s].to(dev, non_blocking=True)                      # H2D of this step's inputs (pinned)
        out = step(m)
        frame_host.copy_(out[0], non_blocking=True)                      # D2H of the frame this rank produced
        torch.cuda.current_stream().synchronize()

    # ---- warm-up (also builds the CUDA graph)
    pyr.clear()
    for s in range(args.warmup):
        resident_step(s)
    torch.cuda.synchronize()
    # ours: rasterize (one launch per view) + resolve/gather + the net; N > 1 adds the level-0 clear (NCCL's kernel not counted)
    launches_per_step = (2 + eng.n_launches()) if world == 1 else (2 + B + eng.n_launches())
    sampler = ClockSampler(local) if rank == 0 else None
    ms_res = timed(resident_step, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    for s in range(min(3, args.warmup)):
        e2e_step(s)
    ms_e2e = timed(e2e_step, args.steps, args.warmup)
    fps = B * args.steps / (ms_res * 1e-3)
    fps_e2e = B * args.steps / (ms_e2e * 1e-3)

    # ---- live per-kernel measurements for the rooflines (CUDA events on the launching stream, eager launches)
    def time_call(fn, reps=5):
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts[1:]))

    sp = L.stream_ptr()
    tc_ms = tc_flops = gen_ms = gen_flops = tcg_ms = tcg_flops = tco_ms = tco_flops = 0.0
    layer_rows = []
    aux_ms = 0.0
    for ly in eng.ops:
        t = time_call(lambda ly=ly: eng.launch_op(ly, sp), reps=4)
        if ly.plan is None:
            aux_ms += t
            layer_rows.append({"name": ly.name, "impl": -1, "ms": t, "gflop": 0.0, "tflops": 0.0})
            continue
        layer_rows.append({"name": ly.name, "impl": int(ly.impl), "ms": t, "gflop": ly.flops / 1e9,
                           "tflops": ly.flops / (t * 1e-3) / 1e12})
        if ly.impl == L.CONV_TCGEN05 and ly.k == 3 and ly.stride == 1:
            tc_ms += t; tc_flops += ly.flops          # the tensor-bound instances: 3x3 stride-1 C->C convs
        elif ly.impl == L.CONV_TCGEN05:
            tco_ms += t; tco_flops += ly.flops        # 1x1 / stride-2 / RAW-term launches: HBM- and latency-bound
        elif ly.impl == L.CONV_TCGEN05_GATHER:
            tcg_ms += t; tcg_flops += ly.flops
        else:
            gen_ms += t; gen_flops += ly.flops
    m0 = mats_dev[args.warmup]

    def project(m):
        ops.raster_project_sorted(pyr, store, m)

    def raster_frame():                      # what a frame does before the net (level 0 is clean on entry)
        project(m0)
        ops.pyramid_resolve_gather(tex_nd, pyr, eng.inputs, layout, view0=rank if world > 1 else 0,
                                   nviews=1 if world > 1 else None, reset_level0=(world == 1))
        if world > 1:
            L.check(lib.read_zbuf_clear(pyr.buf.data_ptr(), pyr.B * W * H, sp))

    def one_shot(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    pyr.clear()
    rg_ms = time_call(raster_frame, reps=6)
    # the two halves separately (each on the state the other leaves behind)
    project_ms = float(np.mean([one_shot(lambda: project(m0)) +
                                0 * one_shot(lambda: ops.pyramid_resolve_gather(tex_nd, pyr, eng.inputs, layout, view0=0,
                                                                                nviews=1, reset_level0=True))
                                for _ in range(4)][1:]))
    resolve_ms = 0.0
    for _ in range(4):
        project(m0)
        resolve_ms += one_shot(lambda: ops.pyramid_resolve_gather(tex_nd, pyr, eng.inputs, layout, view0=0, nviews=1,
                                                                   reset_level0=True)) / 4
    if world > 1:
        pyr.clear()
    raster_ms, gather_ms = project_ms, resolve_ms
    P = sum(w_l * h_l for (w_l, h_l) in pyr.sizes)
    feat_bytes = 2 if eng.bf16 else 4
    # algorithmic bytes: xyz once + packed z write + descriptor read + feature write
    raster_bytes = 12 * count * 1 + P * B * 8
    gather_bytes = P * (8 + 32 + 8 * feat_bytes)
    rg_bytes = raster_bytes + gather_bytes
    hbm = pk["hbm_gbs"]
    tens_peak = pk["bf16_tflops_sustained"]
    roof_tc = None
    traf = measured_traffic()
    if tc_ms > 0:
        ach = tc_flops / (tc_ms * 1e-3) / 1e12
        roof_tc = {"kernel": "gated_conv_tc_kernel<3,*,*,*,*,1> (tcgen05 implicit-GEMM gated conv, 3x3 stride-1 instances = "
                             f"{100.0 * tc_flops / max(eng.flops, 1):.1f}% of the net's conv FLOPs)", "bound": "tensor",
                   "achieved": ach, "peak": tens_peak, "unit": "TFLOP/s", "frac": ach / tens_peak,
                   "peak_src": pk["src"] + " (sustained bf16)",
                   "traffic": (traf or {}).get("gated_conv_tc_kernel_avg_bytes_per_launch"),
                   "traffic_src": "profiles/r01_traffic.json (ncu dram bytes, average over the 3x3 launches)" if traf else None,
                   "ms_per_frame": tc_ms,
                   "layers": sum(1 for l_ in eng.layers if l_.impl == L.CONV_TCGEN05 and l_.k == 3 and l_.stride == 1)}
    ach_r = rg_bytes / (rg_ms * 1e-3) / 1e9
    roof_raster = {"kernel": "raster_sorted_kernel + pyramid_resolve_gather_kernel",
                   "bound": "hbm",
                   "achieved": ach_r, "peak": hbm, "unit": "GB/s", "frac": ach_r / hbm, "peak_src": pk["src"],
                   "traffic": ((traf.get("raster_sorted_kernel_bytes_per_launch", traf["raster_lean_kernel_bytes_per_launch"])
                                + traf["pyramid_resolve_gather_bytes_per_launch"]) if traf else None),
                   "algorithmic_bytes": rg_bytes, "ms_per_frame": rg_ms,
                   "note": "algorithmic bytes count 12 B per point (SURVEY 8d); the sorted store holds 16 B per point (xyz + original id)",
                   "project_ms": project_ms, "resolve_gather_ms": resolve_ms}
    gen_ach = gen_flops / (gen_ms * 1e-3) / 1e12 if gen_ms > 0 else None
    tcg_ach = tcg_flops / (tcg_ms * 1e-3) / 1e12 if tcg_ms > 0 else None

    if rank == 0 and args.layer_times:
        os.makedirs(os.path.dirname(os.path.abspath(args.layer_times)), exist_ok=True)
        json.dump(layer_rows, open(args.layer_times, "w"), indent=0)
    if rank == 0:
        import torch as _t
        cpu_line = None
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            oracle.build()
            sd_cpu = synth.synth_state_dict(synth.SEED)
            cores, avail = pick_torch_threads(sd_cpu)
            t_frame, info = cpu_reference_frame(xyz_np, tex.texture_.detach().cpu(), sd_cpu, cores)
            cpu_line = {"value": 1.0 / t_frame, "unit": "frames/s", "cores": cores, "kind": "port",
                        "sample": (f"1 frame: full-size sequential z-buffer (4 levels, 10M pts, 1 thread/level) = "
                                   f"{info['raster_s']:.2f}s + gather+net on a 1024x512 window = {info['net_crop_s']:.2f}s "
                                   f"x{info['scale']:.2f} by pixel count; torch threads {cores} (fastest of a sweep, {avail} available)")}
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if eng.bf16 else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_points": N_POINTS, "width": W, "height": H, "levels": LEVELS,
                       "views_per_step": B, "parallelism": f"spatial-tile point shards x{world} (Morton ranges) + one NCCL min-reduce + frame-parallel net" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (120 MB cloud, 16.7 MB z-buffer, >130 MB activations per layer at full res)",
                       "cuda_graph": bool(eng.use_graph), "conv_impl": eng.impl_histogram()},
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(B * 64),
                    "d2h_bytes_per_step": int(3 * H * W * 4), "ms_per_step": ms_e2e / args.steps,
                    "note": "camera matrices from pinned host memory per step; point cloud/descriptors/weights are scene state resident in HBM (as MyRender.update_ds / load_textures); RGB frame copied to pinned host memory and synchronised every step"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "roofline": roof_tc if roof_tc else roof_raster,
            "roofline_raster": roof_raster,
            "breakdown_ms_per_frame": {"raster_project": raster_ms, "pyramid_resolve_gather": gather_ms, "raster_total": rg_ms,
                                       "conv_tcgen05_tma_3x3": tc_ms, "conv_tcgen05_tma_other": tco_ms,
                                       "conv_tcgen05_tma_other_tflops": (tco_flops / (tco_ms * 1e-3) / 1e12 if tco_ms > 0 else None),
                                       "conv_tcgen05_gather": tcg_ms,
                                       "conv_tcgen05_gather_tflops": tcg_ach, "conv_generic": gen_ms, "upsample_kernels": aux_ms,
                                       "conv_generic_tflops": gen_ach, "net_flops": eng.flops},
            "cpu_baseline": cpu_line,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layer-times", default=None, help="write per-layer CUDA-event timings (JSON) to this path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
