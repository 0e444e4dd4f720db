"""torchrun script: the fused coupling + peer-memory exchange against the NCCL allgather (2-8 GPUs).
With MSGPU_P2P_SAME_DEVICE=1 all ranks use cuda:0 and no NCCL communicator (one-GPU box)."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
same = os.environ.get("MSGPU_P2P_SAME_DEVICE") == "1"
dev = 0 if same else local
torch.cuda.set_device(dev)
dist.init_process_group("gloo" if same else "nccl", **({} if same else {"device_id": torch.device("cuda", dev)}))
from dealiiscale_b200 import capi
from dealiiscale_b200.micro import MicroSolver, unique_id
from dealiiscale_b200.prm import BioData

macro = int(sys.argv[1]) if len(sys.argv) > 1 else 12
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 16
g = np.linspace(-1, 1, macro + 1)
xy = np.array([(a, b) for b in g for a in g])
n_total = len(xy)
k0, k1 = capi.partition(n_total, world, rank)
ms = MicroSolver(capi.BIO, BioData(os.path.join(ROOT, "input", "biocase.prm")), cells, device=dev)
ms.set_grid_locations(xy[k0:k1]); ms.setup()
uid = [unique_id() if rank == 0 and not same else None]
dist.broadcast_object_list(uid, src=0)
ms.attach_communicator(uid[0], rank, world, n_total)
u, w = xy[:, 1] * np.sin(xy[:, 0]), xy[:, 0] * np.cos(xy[:, 1])

def cycle(scale):
    ms.set_macro_solution(scale * u[k0:k1], scale * w[k0:k1])
    ms.run()
    return ms.coupling()

ref = []
if not same:
    ref = [cycle(s) for s in (1.0, 0.5, 2.0)]
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(50):
        ms.coupling_on_device()
    ms.synchronize(); t_nccl = (time.perf_counter() - t0) / 50
handles = [None] * world
dist.all_gather_object(handles, ms.p2p_export())
ms.p2p_attach(handles)
ms.zero_solutions()
got = [cycle(s) for s in (1.0, 0.5, 2.0)]
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(50):
    ms.coupling_on_device()
ms.synchronize(); t_p2p = (time.perf_counter() - t0) / 50
ok = True
if same:
    # reference: one context with all grids
    if rank == 0:
        one = MicroSolver(capi.BIO, BioData(os.path.join(ROOT, "input", "biocase.prm")), cells, device=dev)
        one.set_grid_locations(xy); one.setup()
        for i, s in enumerate((1.0, 0.5, 2.0)):
            one.set_macro_solution(s * u, s * w); one.run()
            cu, cw = one.coupling()
            ok &= np.array_equal(cu, got[i][0]) and np.array_equal(cw, got[i][1])
        one.close()
else:
    for a, b in zip(ref, got):
        ok &= np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
flag = torch.tensor([1 if ok else 0], device=None if same else "cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"ranks {world} N {n_total}: peer-memory exchange equals reference: {bool(flag.item())}; coupling call "
          f"{'nccl %.1f us, ' % (t_nccl * 1e6) if not same else ''}p2p {t_p2p * 1e6:.1f} us", flush=True)
ms.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
