"""torchrun worker: per-stage CUDA-event times of the distributed CH step (p2p transport)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from evoxels_b200.distributed import DistributedCahnHilliardIMEX
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import weak_scaling_shape
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = weak_scaling_shape(512, world)
st = DistributedCahnHilliardIMEX(shape, (1.0, 1.0, 1.0), 0.1, device=dev, transport="p2p", overlap_chunks=1)
u = 0.5 + 0.1 * torch.rand(st.slab.local_shape, device=dev)
for _ in range(5): u = st.step(u)
acc = {}
for _ in range(10):
    dist.barrier(); torch.cuda.synchronize()
    u, t = st.step_profiled(u)
    for k, v in t.items(): acc[k] = acc.get(k, 0.0) + v / 10
if rank == 0:
    print(json.dumps({"world": world, "shape": shape, "stages_ms": {k: round(v, 3) for k, v in acc.items()}, "sum": round(sum(acc.values()), 3)}))
dist.destroy_process_group()
