"""torchrun worker: distributed (x-slab) steps on N GPUs vs the single-GPU step.
    torchrun --nproc-per-node N tests/dist_gpu_worker.py
Every rank also computes the full-domain single-GPU result itself (small grids) and compares
its slab.  Prints 'DIST-GPU OK' on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import evoxels_b200 as evo  # noqa: E402
from evoxels_b200.distributed import DistributedAllenCahnEuler, DistributedCahnHilliardIMEX, Slab  # noqa: E402
from evoxels_b200.problem_definition import CahnHilliard, TwoPhaseAllenCahn  # noqa: E402
from evoxels_b200.timesteppers import ForwardEuler, PseudoSpectralIMEX  # noqa: E402
from evoxels_b200.voxelgrid import VoxelGridTorch  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    worst = 0.0
    for shape in [(64, 64, 64), (256, 128, 64), (32, 64, 256)]:
        if shape[0] % world or shape[1] % world:
            continue
        spacing = (1.0, 0.5, 2.0)
        gen = torch.Generator(device=dev).manual_seed(7)
        u = 0.5 + 0.1 * torch.rand(shape, device=dev, generator=gen)
        dom = tuple(float(n * h) for n, h in zip(shape, spacing))
        vf = evo.VoxelFields(shape, dom)
        vg = VoxelGridTorch(vf.grid_info(), device=str(dev))
        single = PseudoSpectralIMEX(CahnHilliard(vg), 0.1)
        slab = Slab(shape, world, rank)
        results = {}
        for transport in ("nccl", "p2p", "ce"):
            stepper = DistributedCahnHilliardIMEX(shape, spacing, 0.1, device=dev, transport=transport)
            v, w = u[None], slab.take(u).contiguous()
            m0 = stepper.total_mass(w)
            for _ in range(3):
                v, w = single.step(0.0, v), stepper.step(w)
            ref = slab.take(v[0])
            err = float((w - ref).norm() / ref.norm())
            upd = float(((w - slab.take(u)) - (ref - slab.take(u))).norm() / (ref - slab.take(u)).norm())
            assert err < 1e-6 and upd < 1e-5, (shape, transport, rank, err, upd)
            assert abs(stepper.total_mass(w) - m0) <= 2e-7 * abs(m0)
            worst = max(worst, upd)
            results[transport] = w
        assert torch.equal(results["nccl"], results["p2p"]), (shape, rank)   # same arithmetic
        assert torch.equal(results["nccl"], results["ce"]), (shape, rank)
        for bc in (("neumann",) * 3, ("periodic",) * 3, (("dirichlet", (0.0, 1.0)), "neumann", "periodic")):
            phi = torch.rand(shape, device=dev, generator=gen)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                prob = TwoPhaseAllenCahn(vg, bc=bc)
            euler = ForwardEuler(prob, 0.05)
            ac = DistributedAllenCahnEuler(shape, spacing, 0.05, bc=bc, device=dev)
            v, w = phi[None], slab.take(phi).contiguous()
            for _ in range(2):
                v, w = euler.step(0.0, v), ac.step(w)
            assert torch.equal(w, slab.take(v[0])), (shape, bc, rank)   # same kernel, same inputs
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"DIST-GPU OK world={world} worst update rel-L2 {float(t):.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
