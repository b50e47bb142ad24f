"""What compute-sanitizer runs: smoke() and one C2 frame (1M points, 512x512) through the viewer plugin object,
eager launches (no CUDA graph: the sanitizer patches individual launches).
    compute-sanitizer --tool memcheck python scripts/sanitize_frame.py [smoke|c2|both]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

what = sys.argv[1] if len(sys.argv) > 1 else "both"
if what in ("smoke", "both"):
    import __graft_entry__ as g
    g.smoke()
    print("smoke ok", flush=True)
if what in ("c2", "both"):
    from read_b200 import synth, _lib as L
    from read_b200.viewer import FrameRenderer
    W = H = 512
    n = 1_000_000
    dev = torch.device("cuda", 0)
    xyz = synth.street_scene(n, depth=250.0)
    sd = synth.synth_state_dict(synth.SEED)
    tex = torch.rand((1, 8, n), generator=torch.Generator().manual_seed(1))
    fr = FrameRenderer(xyz, sd, tex, (W, H), device=dev)
    fr.model.net.use_graph = False
    for pose in (3, 7):
        proj, view = synth.camera_batch(W, H, [pose])
        out = fr.infer(proj[0], view[0])
        torch.cuda.synchronize()
        o = out["output"]
        assert torch.isfinite(o).all()
    print("c2 frames ok", tuple(o.shape), float(o.float().mean()), flush=True)
