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


def api_drop_in(dev, rank, world):
    """The reference's own entry points under torchrun: run_cahn_hilliard_solver /
    run_allen_cahn_solver / a user potential decompose the grid over the ranks by themselves and
    must leave the SAME global fields in `vf` on every rank as a one-GPU run of the same call."""
    import numpy as np
    from evoxels_b200.solvers import TimeDependentSolver
    shape = (64, 32, 64)         # the distributed FFT plan wants power-of-two extents
    rng = np.random.default_rng(5)
    c0 = (0.5 + 0.1 * rng.random(shape)).astype(np.float32)
    mu = lambda c, lib=None: 4.0 * c * (1 - c) * (1 - 2 * c) + 0.3 * c      # noqa: E731

    def fields(distributed, problem_cls, stepper_cls, kw, dt):
        vf = evo.VoxelFields(shape, tuple(float(n) for n in shape))
        vf.add_field("c", c0.copy())
        s = TimeDependentSolver(vf, "c", "torch", problem_cls=problem_cls, timestepper_cls=stepper_cls,
                                device=str(dev), distributed=distributed)
        s.solve(time_increment=dt, frames=2, max_iters=6, problem_kwargs=kw, jit=True, verbose=False)
        assert (s.vg.slab is not None) == bool(distributed)
        return vf.fields["c"]

    for cls, ts, kw, dt, tol in ((CahnHilliard, PseudoSpectralIMEX, dict(eps=3.0, D=1.0), 0.1, 1e-6),
                                 (CahnHilliard, PseudoSpectralIMEX, dict(eps=3.0, D=1.0, mu_hom=mu), 0.1, 1e-6),
                                 (TwoPhaseAllenCahn, ForwardEuler, dict(), 0.05, 0.0)):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            one = fields(False, cls, ts, kw, dt)
            many = fields("auto", cls, ts, kw, dt)
        assert one.shape == many.shape == shape
        err = float(np.linalg.norm(many - one) / np.linalg.norm(one))
        assert err <= tol, (cls.__name__, kw.keys(), rank, err)
    # the precompiled driver with the reference's signature
    vf = evo.VoxelFields(shape, tuple(float(n) for n in shape))
    vf.add_field("c", c0.copy())
    solver = evo.run_cahn_hilliard_solver(vf, "c", backend="torch", device=str(dev), time_increment=0.1,
                                          frames=2, max_iters=4, verbose=False)
    assert solver.vg.slab is not None and vf.fields["c"].shape == shape and np.isfinite(vf.fields["c"]).all()


def adjoint(dev, rank, world):
    """Backprop through two distributed steps (last row): the slab of dL/du and the
    all-reduced dL/dD, dL/deps against the single-GPU hand-written adjoint on the whole field."""
    shape, spacing = (64, 32, 128), (1.0, 0.5, 2.0)
    if shape[0] % world or shape[1] % world or shape[0] // world < DistributedCahnHilliardIMEX.ADJ_HALO:
        return
    gen = torch.Generator(device=dev).manual_seed(3)
    u0 = 0.5 + 0.3 * torch.rand(shape, device=dev, generator=gen)
    tgt = 0.5 + 0.1 * torch.rand(shape, device=dev, generator=gen)
    dom = tuple(float(n * h) for n, h in zip(shape, spacing))
    vf = evo.VoxelFields(shape, dom)
    vg = VoxelGridTorch(vf.grid_info(), device=str(dev), distributed=False)
    D1 = torch.tensor(1.3, dtype=torch.float64, device=dev, requires_grad=True)
    e1 = torch.tensor(2.5, dtype=torch.float64, device=dev, requires_grad=True)
    single = PseudoSpectralIMEX(CahnHilliard(vg, eps=e1, D=D1), 0.1)
    v = u0[None].clone().requires_grad_(True)
    x = v
    for _ in range(2):
        x = single.step(0.0, x)
    ((x[0] - tgt) ** 2).sum().backward()
    slab = Slab(shape, world, rank)
    for transport in ("nccl", "ce"):
        D2 = torch.tensor(1.3, dtype=torch.float64, device=dev, requires_grad=True)
        e2 = torch.tensor(2.5, dtype=torch.float64, device=dev, requires_grad=True)
        st = DistributedCahnHilliardIMEX(shape, spacing, 0.1, eps=2.5, D=1.3, device=dev, transport=transport)
        w = slab.take(u0).contiguous().requires_grad_(True)
        y = w
        for _ in range(2):
            y = st.step_autograd(y, D2, e2)
        ((y - slab.take(tgt)) ** 2).sum().backward()
        ref = slab.take(v.grad[0])
        err = float((w.grad - ref).norm() / ref.norm())
        assert err < 1e-5, ("adjoint dL/du", transport, rank, err)
        assert abs(float(D2.grad) - float(D1.grad)) <= 1e-5 * abs(float(D1.grad)), (float(D2.grad), float(D1.grad))
        assert abs(float(e2.grad) - float(e1.grad)) <= 1e-5 * abs(float(e1.grad)), (float(e2.grad), float(e1.grad))


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
        vg = VoxelGridTorch(vf.grid_info(), device=str(dev), distributed=False)
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
    api_drop_in(dev, rank, world)
    adjoint(dev, rank, world)
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"DIST-GPU OK world={world} worst update rel-L2 {float(t):.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
