"""torchrun worker: A/B of the transports of the distributed CH step on real GPUs (ms/step, max over
ranks, 20 steps after 5), all variants inside one session; every variant is checked against the
first one (bit-identical fields after the same number of steps)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
from evoxels_b200.distributed import DistributedCahnHilliardIMEX, Slab

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 512
shape = {1: (n, n, n), 2: (2 * n, n, n), 4: (2 * n, 2 * n, n), 8: (2 * n, 2 * n, 2 * n)}[world]
if len(sys.argv) > 1 and "x" in sys.argv[1]:
    shape = tuple(int(v) for v in sys.argv[1].split("x"))
variants = [
    ("ce4", dict(transport="ce", overlap_chunks=4), {}),
    ("ce4_batch_fwd", dict(transport="ce", overlap_chunks=4), {"batch": 2}),
    ("ce4_batch_all", dict(transport="ce", overlap_chunks=4), {"batch": 1}),
    ("ce4_batch_all_lastdma", dict(transport="ce", overlap_chunks=4), {"batch": 1, "EVX_CE_LAST_CTAS": "0"}),
    ("ce4_again", dict(transport="ce", overlap_chunks=4), {}),
]
only = os.environ.get("EVX_AB_ONLY")
if only:
    variants = [v for v in variants if v[0] in only.split(",")]
gen = torch.Generator(device="cuda").manual_seed(100 + rank)
u0 = 0.5 + 0.1 * torch.rand(Slab(shape, world, rank).local_shape, device=dev, generator=gen)
out = {}
ref = None
for name, kw, env in variants:
    attrs = {k: v for k, v in env.items() if not k.startswith("EVX_")}
    env = {k: v for k, v in env.items() if k.startswith("EVX_")}
    for k, v in env.items():
        os.environ[k] = v
    st = DistributedCahnHilliardIMEX(shape, (1.0, 1.0, 1.0), 0.1, device=dev, **kw)
    if "EVX_CE_LAST_CTAS" in env:
        st.ops.last_chunk_ctas = int(env["EVX_CE_LAST_CTAS"])
    if "batch" in attrs:
        st.ops.batch_copies = attrs["batch"]
    if "direct" in attrs:
        st.ops.direct_peers, st.ops.direct_peers_mid = attrs["direct"], attrs["direct_mid"]
    u = u0.clone()
    try:
        for _ in range(5):
            u = st.step(u)
        torch.cuda.synchronize()
    except Exception as exc:                      # e.g. an API the driver refuses: same on every rank
        if rank == 0:
            print(name, "FAILED", repr(exc)[:300], flush=True)
        out[name] = {"error": repr(exc)[:300]}
        del st
        torch.cuda.empty_cache()
        dist.barrier()
        continue
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        u = st.step(u)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if ref is None:
        ref = u.clone()
    same = torch.tensor([int(torch.equal(u, ref))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out[name] = {"ms_per_step": round(float(ms.item()), 4), "bit_equal_to_first": bool(same.item()),
                 "finite": bool(torch.isfinite(u).all().item())}
    if rank == 0:
        print(name, json.dumps(out[name]), flush=True)
    for k in env:
        os.environ.pop(k, None)
    del st
    torch.cuda.empty_cache()
    dist.barrier()
if rank == 0:
    os.makedirs("gpurun_out/r02_ab", exist_ok=True)
    json.dump({"shape": shape, "world": world, "variants": out}, open(f"gpurun_out/r02_ab/ab_n{world}.json", "w"), indent=1)
dist.destroy_process_group()
