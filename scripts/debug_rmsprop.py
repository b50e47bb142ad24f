import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from read_b200 import train
from read_b200.texture import PointTexture
dev = torch.device("cuda", 0)
N, D = 20000, 8
gen = torch.Generator().manual_seed(1)
tex = PointTexture(D, N, init_method='rand').to(dev)
ref = torch.nn.Parameter(tex.texture_.detach().clone())
opt_ref = torch.optim.RMSprop([ref], lr=0.1)
opt = train.SparseRMSprop(tex, lr=0.1)
for step in range(8):
    pool = torch.randperm(N - 1, generator=gen)[: 300 + 700 * (step % 3)] + 1
    ids = pool[torch.randint(0, len(pool), (2, 1, 32, 32), generator=gen)].float()
    ids[torch.rand((2, 1, 32, 32), generator=gen) < 0.3] = 0.
    ids = ids.to(dev)
    up = torch.randn((2, D, 32, 32), generator=gen).to(dev)
    opt_ref.zero_grad()
    idx = ids[:, 0].long().reshape(-1)
    smp = torch.index_select(ref[0], 1, idx).view(D, 2, 32, 32).permute(1, 0, 2, 3)
    (smp * up).sum().backward()
    (tex(ids) * up).sum().backward()
    g_ours = tex._sparse.grad.t().clone()
    gerr = float((g_ours - ref.grad[0]).abs().max())
    tset = tex._sparse.touched.bool().clone()
    want_t = (ref.grad[0].abs().sum(0) > 0)
    opt_ref.step(); opt.step()
    torch.cuda.synchronize()
    d = (tex.texture_.detach() - ref.detach()).abs()[0]
    bad = (d.max(0).values > 1e-4).nonzero().flatten()
    print(f"step {step}: grad err {gerr:.2e}, touched {int(tset.sum())} want {int(want_t.sum())} flag-mismatch {int((tset != want_t).sum())}, param err {float(d.max()):.3e}, bad points {bad[:8].tolist()} ({len(bad)})")
    if len(bad):
        i = int(bad[0])
        print("   point", i, "ours", tex.texture_[0, :, i].tolist()[:3], "ref", ref[0, :, i].tolist()[:3], "touched", bool(tset[i]), "grad", ref.grad[0][:3, i].tolist())
