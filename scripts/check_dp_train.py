"""torchrun check of the data-parallel descriptor-gradient join: after exchange_sparse_grads every rank holds the SUM of all
ranks' touched gradient rows (== all-reducing the dense [N,D] gradient), and SparseRMSprop keeps the replicas identical.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_dp_train.py"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import train                       # noqa: E402
from read_b200.texture import PointTexture        # noqa: E402
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, D = 200_000, 8
tex = PointTexture(D, N, init_method='zeros').to(dev)
with torch.no_grad():
    tex.texture_.copy_(torch.rand((1, D, N), generator=torch.Generator().manual_seed(0)).to(dev))
opt = train.SparseRMSprop(tex, lr=0.1)
ok = True
for step in range(4):
    g = torch.Generator().manual_seed(100 * step + rank)
    ids = torch.randint(0, N, (4, 1, 64, 64), generator=g).float().to(dev)
    up = torch.randn((4, D, 64, 64), generator=g).to(dev)
    (tex(ids) * up).sum().backward()
    mine = tex._sparse.grad.clone()
    dense = mine.clone()
    dist.all_reduce(dense)                              # what a dense gradient all-reduce would give
    train.exchange_sparse_grads(tex)
    err = float((tex._sparse.grad - dense).abs().max())
    t_union = (dense.abs().sum(1) > 0)
    flags_ok = bool(torch.equal(tex._sparse.touched.bool() | ~t_union, torch.ones_like(t_union)))
    opt.step()
    chk = tex.texture_.detach().double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok = ok and err < 1e-4 and flags_ok and float(hi - lo) == 0.0
    if rank == 0:
        print(f"step {step}: max |sparse-joined - dense all-reduce| = {err:.2e}, flags cover the union: {flags_ok}, replicas identical: {float(hi - lo) == 0.0}")
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("data-parallel descriptor join OK:", bool(flag.item()))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
