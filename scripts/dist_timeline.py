"""torchrun worker: event timeline of one distributed CH step with the copy-engine transport."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from evoxels_b200.distributed import DistributedCahnHilliardIMEX, Slab

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
shape = (n * world, n, n) if world <= 2 else ((n * 2, n * 2, n * world // 4) if world == 8 else (n * 2, n * 2, n))
if len(sys.argv) > 3:
    shape = tuple(int(v) for v in sys.argv[3].split("x"))
st = DistributedCahnHilliardIMEX(shape, (1.0, 1.0, 1.0), 0.1, device=dev, transport="ce", overlap_chunks=chunks,
                                 copier=os.environ.get("EVX_COPIER"), scatter_ctas=int(os.environ.get("EVX_SCATTER_CTAS", "0")))
u = 0.5 + 0.1 * torch.rand(Slab(shape, world, rank).local_shape, device=dev)
for _ in range(5):
    u = st.step(u)
torch.cuda.synchronize(); dist.barrier()
st.ops.trace = []
t0 = torch.cuda.Event(enable_timing=True); t0.record()
u = st.step(u)
torch.cuda.synchronize()
if rank == 0:
    print(f"shape {shape} world {world} chunks {chunks}")
    for name, e in sorted(st.ops.trace, key=lambda p: t0.elapsed_time(p[1])):
        print(f"{t0.elapsed_time(e):8.3f} ms  {name}")
dist.destroy_process_group()
