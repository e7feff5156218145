"""Worker for tests/test_distributed_cpu.py: runs the distributed orchestration of
evoxels_b200.distributed on `world` CPU processes (gloo) with a torch-CPU stand-in for the
kernels, and compares every rank's slab with the single-domain oracle.

The stand-in (`OracleOps`) is TEST CODE: it reproduces what the CUDA kernels compute for
one rank (rhs from slab + halo planes; z/y transforms into the all-to-all block layout; x
transform + filter with the global ky offset; inverse) using the oracle's arithmetic, so
that slab geometry, halo routing, block ordering and global index offsets are verified
without a GPU.  The shipped CudaOps has no CPU path.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import evx_oracle as O  # noqa: E402
from evoxels_b200.distributed import (DistributedAllenCahnEuler, DistributedCahnHilliardIMEX,  # noqa: E402
                                      Slab)


class OracleOps:
    def __init__(self, slab, spacing, dt, coef_power):
        self.slab, self.spacing = slab, tuple(spacing)
        nx, ny, nz = slab.global_shape
        self.nzh = nz // 2 + 1
        shape = (slab.world, slab.nxl, slab.nyl, self.nzh)
        self.a = torch.zeros(shape, dtype=torch.complex64)
        self.b = torch.zeros(shape, dtype=torch.complex64)

    def new_field(self):
        return torch.empty(self.slab.local_shape, dtype=torch.float32)

    @staticmethod
    def _extended(u, lo, hi, width):
        parts = ([lo] if lo is not None else []) + [u] + ([hi] if hi is not None else [])
        return torch.cat(parts, 0), (width if lo is not None else 0), (width if hi is not None else 0)

    def ch_rhs(self, u, out, eps, D, bc, halo_lo, halo_hi):
        ext, a, b = self._extended(u, halo_lo, halo_hi, 2)
        # x rule of `bc` applies only where no halo was supplied (domain ends); with halos on
        # both sides any x rule is fine because the footprint (radius 2) stays inside `ext`
        r = O.ch_rhs(ext[None], self.spacing, eps, D, bc)[0]
        out.copy_(r[a:r.shape[0] - b])

    user_mu = None

    def ch_rhs_hom(self, c, out, eps, D, bc, hom):
        # the stand-in re-evaluates the potential inside the oracle (same function the test
        # handed to the stepper as hom_fn) and checks that the caller's field is that function
        assert torch.allclose(hom, self.user_mu(torch.clip(c, 0, 1)))
        out.copy_(O.ch_rhs(c[None], self.spacing, eps, D, bc, self.user_mu)[0])

    def ac_stage(self, phi, out, params, bc, dt, halo_lo, halo_hi):
        ext, a, b = self._extended(phi, halo_lo, halo_hi, 1)
        kw = dict(params)
        r = O.ac_rhs(ext[None], self.spacing, bc=bc, **kw)[0]
        out.copy_(phi + dt * r[a:r.shape[0] - b])

    def exchange_buffers(self):
        return self.a, self.b

    def spectral_forward(self, r):
        s = torch.fft.fft(torch.fft.rfft(r, dim=2), dim=1)            # [nxl, ny, nzh]
        nyl = self.slab.nyl
        for j in range(self.slab.world):
            self.a[j] = s[:, j * nyl:(j + 1) * nyl, :]
        return self.a

    def spectral_middle(self, buf, dt, coef, power):
        nx, ny, nz = self.slab.global_shape
        pencils = buf.reshape(nx, self.slab.nyl, self.nzh)
        k2 = O.k_squared((nx, ny, nz), self.spacing)
        pref = O.imex_prefactor(-coef * k2 ** power, dt)
        y0 = self.slab.rank * self.slab.nyl
        f = torch.fft.fft(pencils, dim=0) * pref[:, y0:y0 + self.slab.nyl, :]
        buf.copy_(torch.fft.ifft(f, dim=0).reshape(buf.shape))

    def spectral_backward(self, buf, u, out):
        s = torch.cat([buf[j] for j in range(self.slab.world)], dim=1)   # [nxl, ny, nzh]
        upd = torch.fft.irfft(torch.fft.ifft(s, dim=1), n=self.slab.global_shape[2], dim=2)
        out.copy_(upd if u is None else u + upd)

    def ch_rhs_vjp_ext(self, u_ext, w_ext, lam_in_ext, eps, D, x_lo, x_hi):
        """Stand-in for the adjoint stencil kernels: autograd through the oracle's rhs on the
        extended slab (float64).  dL/du takes the full cotangent; the dL/deps partial sum takes
        the cotangent of the slab's own planes only (partition by rhs site: the ranks' partial
        sums add up to the global value just like the kernels' partition by mu site)."""
        per = ("periodic",) * 3
        with torch.enable_grad():        # called from inside an autograd backward
            u = u_ext.double().clone().requires_grad_(True)
            e = torch.tensor(float(eps), dtype=torch.float64, requires_grad=True)
            R = O.ch_rhs(u[None], self.spacing, e, float(D), per)[0]
            wd = w_ext.double()
            lam, = torch.autograd.grad((R * wd).sum(), u, retain_graph=True)
            mask = torch.zeros_like(wd)
            mask[x_lo:x_hi] = 1.0
            deps, = torch.autograd.grad((R * wd * mask).sum(), e)
        return (lam + lam_in_ext.double()).float(), deps


def run(rank, world, port, shape):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    spacing = (1.0, 0.5, 2.0)
    slab = Slab(shape, world, rank)
    u = O.noise_field(shape, seed=3, lo=-0.1, amp=1.2)[0]

    # Cahn-Hilliard IMEX, 3 steps
    ops = OracleOps(slab, spacing, 0.1, (1.5, 2))
    stepper = DistributedCahnHilliardIMEX(shape, spacing, 0.1, ops=ops)
    orc = O.CHOracle(shape, spacing, 0.1)
    v, w = u[None], slab.take(u).clone()
    m0 = stepper.total_mass(w)
    for _ in range(3):
        v, w = orc.step(v), stepper.step(w)
    ref = slab.take(v[0])
    err = float((w - ref).norm() / ref.norm())
    assert err < 2e-6, f"CH rank {rank}: {err}"
    assert abs(stepper.total_mass(w) - m0) < 1e-3 * abs(m0) * 1e-3
    assert abs(m0 - float(u.double().sum())) < 1e-6 * abs(m0)

    # Cahn-Hilliard with a user potential (caller-evaluated field on the halo-extended slab)
    if slab.nxl >= 2:
        mu = lambda c, lib=None: 4.0 * c * (1 - c) * (1 - 2 * c) + 0.3 * c     # noqa: E731
        ops = OracleOps(slab, spacing, 0.1, (1.5, 2))
        ops.user_mu = mu
        stepper = DistributedCahnHilliardIMEX(shape, spacing, 0.1, ops=ops,
                                              hom_fn=lambda c: mu(torch.clip(c, 0, 1)))
        orc = O.CHOracle(shape, spacing, 0.1, mu_hom=mu)
        v, w = u[None], slab.take(u).clone()
        for _ in range(2):
            v, w = orc.step(v), stepper.step(w)
        ref = slab.take(v[0])
        err = float((w - ref).norm() / ref.norm())
        assert err < 2e-6, f"CH user potential rank {rank}: {err}"

    # adjoint of one Cahn-Hilliard step: slab gradient and the all-reduced parameter gradients
    # against float64 autograd through the single-domain oracle
    if world == 1 or slab.nxl >= DistributedCahnHilliardIMEX.ADJ_HALO:
        tgt = O.noise_field(shape, seed=6, lo=0.45, amp=0.1)[0]
        v = u[None].double().clone().requires_grad_(True)
        Dt = torch.tensor(1.3, dtype=torch.float64, requires_grad=True)
        et = torch.tensor(2.5, dtype=torch.float64, requires_grad=True)
        loss = ((O.ch_imex_step(v, spacing, 0.1, et, Dt, 0.25) - tgt[None].double()) ** 2).sum()
        gu, gD, ge = torch.autograd.grad(loss, (v, Dt, et))
        ops = OracleOps(slab, spacing, 0.1, (1.5, 2))
        stepper = DistributedCahnHilliardIMEX(shape, spacing, 0.1, eps=2.5, D=1.3, A=0.25, ops=ops)
        w = slab.take(u).clone().requires_grad_(True)
        Dl = torch.tensor(1.3, dtype=torch.float64, requires_grad=True)
        el = torch.tensor(2.5, dtype=torch.float64, requires_grad=True)
        out = stepper.step_autograd(w, Dl, el)
        local = ((out.double() - slab.take(tgt).double()) ** 2).sum()    # this rank's share of the loss
        local.backward()
        ref = slab.take(gu[0])
        err = float((w.grad.double() - ref).norm() / ref.norm())
        assert err < 2e-4, f"adjoint dL/du rank {rank}: {err}"
        assert abs(float(Dl.grad) - float(gD)) < 2e-4 * abs(float(gD)), (float(Dl.grad), float(gD))
        assert abs(float(el.grad) - float(ge)) < 2e-4 * abs(float(ge)), (float(el.grad), float(ge))

    # Allen-Cahn Euler with three BC layouts along x
    for bc in (("neumann",) * 3, ("periodic",) * 3, (("dirichlet", (0.0, 1.0)), "neumann", "periodic")):
        phi = O.noise_field(shape, seed=1, lo=0.0, amp=1.0)[0]
        ac = DistributedAllenCahnEuler(shape, spacing, 0.05, bc=bc,
                                       ops=OracleOps(slab, spacing, 0.05, (1.0, 1)))
        aorc = O.ACOracle(shape, spacing, 0.05, bc=bc)
        v, w = phi[None], slab.take(phi).clone()
        for _ in range(2):
            v, w = aorc.step(v), ac.step(w)
        ref = slab.take(v[0])
        err = float((w - ref).norm() / ref.norm())
        assert err < 1e-6, f"AC {bc} rank {rank}: {err}"
    # the grid object under a process group: slab in, gathered global field out (host-side
    # bookkeeping only - a CPU grid runs no kernels)
    import evoxels_b200 as evo
    from evoxels_b200.voxelgrid import VoxelGridTorch
    vf = evo.VoxelFields(shape, tuple(float(n) for n in shape))
    vg = VoxelGridTorch(vf.grid_info(), device="cpu")
    a = np.random.default_rng(9).random(shape).astype(np.float32)
    f = vg.init_scalar_field(a)
    if world > 1:
        assert vg.slab is not None and tuple(f.shape) == (1,) + slab.local_shape
        assert np.array_equal(f[0].numpy(), a[slab.x0:slab.x0 + slab.nxl])
    else:
        assert vg.slab is None and tuple(f.shape) == (1,) + tuple(shape)
    assert np.array_equal(vg.export_scalar_field_to_numpy(f), a)
    assert VoxelGridTorch(vf.grid_info(), device="cpu", distributed=False).slab is None
    torch.set_default_device("cpu")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    world = int(sys.argv[1])
    port = int(sys.argv[2])
    shape = tuple(int(x) for x in sys.argv[3].split("x"))
    mp.spawn(run, args=(world, port, shape), nprocs=world, join=True)
    print("OK")
