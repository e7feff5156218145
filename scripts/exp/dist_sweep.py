"""torchrun worker: ms/step of the distributed CH step ('ce' transport) for a few pipeline settings.
   torchrun --nproc-per-node N scripts/exp/dist_sweep.py [n_per_gpu]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from evoxels_b200.distributed import DistributedCahnHilliardIMEX, Slab

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
shape = {1: (n, n, n), 2: (2 * n, n, n), 4: (2 * n, 2 * n, n), 8: (2 * n, 2 * n, 2 * n)}[world]
res = {}
STEPS = 20 if n <= 512 else 8
u0 = 0.5 + 0.1 * torch.rand(Slab(shape, world, rank).local_shape, device=dev)
ref = None
CFGS = ((4, 4, 0), (4, 4, 24), (2, 4, 24), (2, 2, 24), (4, 2, 24), (4, 4, 48), (3, 3, 24), (8, 4, 24))
if len(sys.argv) > 2:
    CFGS = tuple(tuple(int(v) for v in c.split(",")) for c in sys.argv[2:])
for chunks, mid, last in CFGS:
    st = DistributedCahnHilliardIMEX(shape, (1.0, 1.0, 1.0), 0.1, device=dev, transport="ce",
                                     overlap_chunks=chunks, mid_chunks=mid)
    st.ops.last_chunk_ctas = last
    u = u0.clone()
    for _ in range(5):
        u = st.step(u)
    chk = float(u.double().sum())
    if ref is None:
        ref = u.clone()
    same = bool(torch.equal(u, ref))
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(STEPS):
        u = st.step(u)
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / STEPS], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[f"fwd{chunks}_mid{mid}_last{last}"] = (round(float(t), 4), same)
    del st
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps({"shape": shape, "world": world, "ms_per_step": res}))
dist.destroy_process_group()
