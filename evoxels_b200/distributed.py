"""x-slab multi-GPU execution of the hot path (one process per GPU, torch.distributed).

The reference is single-device (no NCCL/MPI code exists upstream); this
module is the multi-GPU form of the same step.  Rank r owns the planes
x in [r*nx/W, (r+1)*nx/W) of every field (dim 0 of the [nx,ny,nz] tensor).

* Stencil halos: 2 raw planes per side for the fused Cahn-Hilliard rhs, 1 for Allen-Cahn,
  exchanged point-to-point with the ring neighbours (NCCL send/recv = NVLink P2P on one
  box) and handed to the kernels as separate `halo_lo` / `halo_hi` buffers.
* Spectral stage: local z and y passes, slab->pencil all-to-all, fused x pass, pencil->slab
  all-to-all, local inverse y and z passes.  The y-pass kernels write / read the
  all-to-all block layout directly (no pack / unpack kernels).

All collective calls go through the small `Comm` adapter so that the orchestration (who
sends which planes where, global wavenumber offsets, block layouts) can be tested on CPU
with the gloo backend and a stand-in for the kernels (tests/test_distributed_cpu.py); the
shipped `CudaOps` has no CPU path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _native
from .problem_definition import normalize_bc


@dataclass
class Slab:
    """Geometry of one rank's x-slab."""
    global_shape: tuple
    world: int
    rank: int

    def __post_init__(self):
        nx, ny, nz = self.global_shape
        if nx % self.world or ny % self.world:
            raise ValueError(f"x-slab decomposition needs world | nx and world | ny, got "
                             f"{self.global_shape} on {self.world} ranks")
        self.nxl = nx // self.world
        self.nyl = ny // self.world
        self.x0 = self.rank * self.nxl
        self.local_shape = (self.nxl, ny, nz)

    def take(self, global_field):
        """Local part of a replicated [.., nx, ny, nz] tensor / array."""
        return global_field[..., self.x0:self.x0 + self.nxl, :, :]


class Comm:
    """Thin adapter over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.backend = dist.get_backend(group) if dist.is_initialized() else "none"

    def exchange_halos(self, u_local, width, periodic):
        """Ring exchange of `width` boundary planes of a [nxl,ny,nz] tensor.  Returns
        (halo_lo, halo_hi): the neighbour planes below / above the slab, or None at a
        non-periodic domain end (or everywhere when world == 1 and the axis is periodic,
        in which case the kernels wrap by index arithmetic)."""
        W, r = self.world, self.rank
        if W == 1:
            return None, None
        if u_local.shape[0] < width:
            raise ValueError("slab thinner than the halo width")
        has_lo = periodic or r > 0
        has_hi = periodic or r < W - 1
        lo, hi = (r - 1) % W, (r + 1) % W
        first = u_local[:width].contiguous()
        last = u_local[-width:].contiguous()
        halo_lo = torch.empty_like(first) if has_lo else None
        halo_hi = torch.empty_like(last) if has_hi else None
        ops = []
        # order matters when lo == hi (W == 2): each side posts "first planes" before "last
        # planes", so the receiver takes the upper neighbour's first planes first
        if has_lo:
            ops.append(dist.P2POp(dist.isend, first, lo, self.group, tag=0))
        if has_hi:
            ops.append(dist.P2POp(dist.isend, last, hi, self.group, tag=1))
        if has_hi:
            ops.append(dist.P2POp(dist.irecv, halo_hi, hi, self.group, tag=0))
        if has_lo:
            ops.append(dist.P2POp(dist.irecv, halo_lo, lo, self.group, tag=1))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        return halo_lo, halo_hi

    def all_to_all_blocks(self, send, recv):
        """Block j of `send` ([W, ...]) goes to rank j; block i of `recv` comes from rank i."""
        if self.world == 1:
            recv.copy_(send)
            return
        if self.backend == "nccl":
            # complex64 is not a NCCL dtype: exchange the float32 view
            s = torch.view_as_real(send) if send.is_complex() else send
            r = torch.view_as_real(recv) if recv.is_complex() else recv
            dist.all_to_all_single(r, s, group=self.group)
            return
        ops = []
        for peer in range(self.world):
            if peer == self.rank:
                recv[peer].copy_(send[peer])
                continue
            ops.append(dist.P2POp(dist.isend, send[peer].contiguous(), peer, self.group))
            ops.append(dist.P2POp(dist.irecv, recv[peer], peer, self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def all_reduce_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return t

    def all_reduce_max(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t


class PeerBuffers:
    """Two all-to-all block buffers per rank, allocated in torch symmetric memory so that
    every rank holds device pointers to every peer's copy (NVLink P2P on one NVSwitch box).
    The transform kernels store their output chunks straight into the owning rank's buffer;
    `barrier()` is the device-side cross-rank barrier enqueued on the current stream."""

    def __init__(self, block_shape, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self._symm = symm_mem
        group = group if group is not None else dist.group.WORLD
        shape = tuple(block_shape) + (2,)                     # complex64 as float32 pairs
        self.raw, self.handles, self.bufs, self.peer_ptrs = [], [], [], []
        for _ in range(2):
            t = symm_mem.empty(*shape, dtype=torch.float32, device=device)
            h = symm_mem.rendezvous(t, group=group)
            self.raw.append(t)
            self.handles.append(h)
            self.bufs.append(torch.view_as_complex(t))
            self.peer_ptrs.append([int(x) for x in h.buffer_ptrs])
        self._channel = 0

    def barrier(self, which):
        self.handles[which].barrier(channel=0)


class PeerHalos:
    """Stencil halos through symmetric memory: every rank owns a [2][width][ny][nz] buffer
    (slot 0 = planes below the slab, slot 1 = planes above); `exchange` has the DMA engines
    write this rank's boundary planes straight into the neighbours' slots and enqueues one
    cross-rank barrier - no NCCL call, no staging copies.  `periodic=False`: the domain ends get
    no halo (None is returned there and the kernels apply the Neumann / Dirichlet ghost rule)."""

    def __init__(self, width, ny, nz, device, group, world, rank):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.width, self.world, self.rank = width, world, rank
        # two sets of slots used alternately: a neighbour may write the halos of step k+1 while
        # this rank's kernel of step k still reads its own (one barrier per step orders a write
        # only against the reads of the step before last)
        self.buf = symm_mem.empty(2, 2, width, ny, nz, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group=group)
        self.ptrs = [int(x) for x in self.handle.buffer_ptrs]
        self.slot_bytes = width * ny * nz * 4
        self.parity = 0

    def exchange(self, u_local, periodic=True):
        W, r, w = self.world, self.rank, self.width
        if u_local.shape[0] < w:
            raise ValueError("slab thinner than the halo width")
        st = torch.cuda.current_stream(u_local.device)
        lo, hi = (r - 1) % W, (r + 1) % W
        has_lo, has_hi = periodic or r > 0, periodic or r < W - 1
        plane = u_local.stride(0) * 4
        par = self.parity
        self.parity ^= 1
        base = par * 2 * self.slot_bytes
        # my first planes are the lower neighbour's "above" halo, my last planes the upper
        # neighbour's "below" halo
        if has_lo:
            _native.copy_async(self.ptrs[lo] + base + self.slot_bytes, u_local.data_ptr(), self.slot_bytes, st)
        if has_hi:
            _native.copy_async(self.ptrs[hi] + base, u_local.data_ptr() + (u_local.shape[0] - w) * plane,
                               self.slot_bytes, st)
        self.handle.barrier(channel=0)
        return (self.buf[par, 0] if has_lo else None), (self.buf[par, 1] if has_hi else None)


class CudaOps:
    """The kernels behind one rank of the distributed step.

    transport = 'p2p': the y pass / x pass write their output straight into the peers'
    buffers over NVLink (fused transform + transpose, no collective call, no pack/unpack);
    transport = 'ce': local block buffers, blocks moved into the peers' buffers by the DMA
    engines (cudaMemcpyAsync on copy streams), pipelined against the next chunk's kernels;
    transport = 'nccl': local block buffers + NCCL all-to-all."""

    def __init__(self, slab: Slab, spacing, device, spectral=True, transport="nccl", group=None):
        self.slab, self.spacing, self.device = slab, tuple(spacing), torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("evoxels_b200 has no CPU path: CudaOps needs a CUDA device")
        self.transport = transport if slab.world > 1 else "nccl"
        self.peers = None
        if spectral:
            self.plan = _native.DistPlan(slab.global_shape, slab.world, slab.rank, device)
            self.spec = self.plan.new_buffer()
            if self.transport in ("p2p", "ce"):
                try:
                    self.peers = PeerBuffers(self.plan.block_shape, self.device, group)
                    if self.transport in ("ce", "p2p"):
                        _, ny, nz = slab.global_shape
                        self.halos = PeerHalos(2, ny, nz, self.device, group, slab.world, slab.rank)
                    self.buf_a, self.buf_b = self.peers.bufs
                except Exception as exc:      # no symmetric memory on this platform / topology
                    import warnings
                    warnings.warn(f"torch symmetric memory is not available ({exc}); the slab<->pencil "
                                  "transposes use NCCL all-to-all instead of peer memory")
                    self.peers, self.transport = None, "nccl"
            if self.transport == "nccl":
                self.buf_a = self.plan.new_buffer()
                self.buf_b = self.plan.new_buffer()

    def new_field(self):
        return torch.empty(self.slab.local_shape, dtype=torch.float32, device=self.device)

    def ch_rhs(self, u, out, eps, D, bc, halo_lo, halo_hi):
        _native.ch_rhs(u, out, self.spacing, eps, D, bc, halo_lo=halo_lo, halo_hi=halo_hi)

    def ch_rhs_hom(self, c, out, eps, D, bc, hom):
        """rhs with a caller-evaluated potential field (no x halos: the caller extends the slab)."""
        _native.ch_rhs(c, out, self.spacing, eps, D, bc, hom=hom)

    def ac_stage(self, phi, out, params, bc, dt, halo_lo, halo_hi):
        _native.ac_stage(phi, self.spacing, params["eps"], params["gab"], params["M"],
                         params["force"], params["curvature"], bc, base=phi, y_out=out, alpha=dt,
                         halo_lo=halo_lo, halo_hi=halo_hi)

    def ch_rhs_vjp_ext(self, u_ext, w_ext, lam_in_ext, eps, D, x_lo, x_hi):
        """Adjoint stencil kernels on a halo-extended slab (periodic inside the kernels: the wrap
        only reaches the outer halo planes, which the caller drops); only the planes
        [x_lo, x_hi) enter the dL/deps partial sum.  Returns (lam_ext, deps)."""
        return _native.ch_rhs_vjp(u_ext, w_ext, self.spacing, eps, D, lam_in=lam_in_ext,
                                  deps_planes=(x_lo, x_hi))

    def spectral_forward(self, r):
        self.plan.forward(r, self.spec, self.buf_a)
        return self.buf_a

    def spectral_middle(self, buf, dt, coef, power):
        self.plan.middle(buf, self.spacing, dt, coef, power)

    # fused transform + transpose over peer memory --------------------------------------------
    def spectral_forward_p2p(self, r):
        """ZFwd + y pass; chunks land in every peer's buffer B.  Returns the local B."""
        self.plan.forward_p2p(r, self.spec, self.peers.peer_ptrs[1])
        self.peers.barrier(1)
        return self.buf_b

    def forward_pipelined(self, u, rhs, eps, D, bc, halo_lo, halo_hi, chunks):
        """rhs -> z pass -> y pass (+ NVLink peer stores) over `chunks` slices of the local x
        range.  A compute stream runs rhs and z pass of slice i+1 while a link stream runs the
        NVLink-bound y pass of slice i (launched with a capped grid so that SMs stay free)."""
        nxl = self.slab.nxl
        bounds = [round(i * nxl / chunks) for i in range(chunks + 1)]
        main = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_comp"):
            self._comp = torch.cuda.Stream(device=self.device)
            self._link = torch.cuda.Stream(device=self.device)
        self._comp.wait_stream(main)
        self._link.wait_stream(main)
        for i in range(chunks):
            x0, x1 = bounds[i], bounds[i + 1]
            with torch.cuda.stream(self._comp):
                lo = halo_lo if x0 == 0 else u[x0 - 2:x0]
                hi = halo_hi if x1 == nxl else u[x1:x1 + 2]
                _native.ch_rhs(u[x0:x1], rhs[x0:x1], self.spacing, eps, D, bc, halo_lo=lo, halo_hi=hi)
                self.plan.forward_chunk_p2p(rhs, self.spec, self.peers.peer_ptrs[1], x0, x1 - x0, parts=1)
                done = torch.cuda.Event()
                done.record(self._comp)
            self._link.wait_event(done)
            with torch.cuda.stream(self._link):
                self.plan.forward_chunk_p2p(rhs, self.spec, self.peers.peer_ptrs[1], x0, x1 - x0, parts=2)
        main.wait_stream(self._comp)
        main.wait_stream(self._link)
        self.peers.barrier(1)
        return self.buf_b

    # optional timeline (diagnostics): set ops.trace = [] to collect (name, event) pairs
    trace = None
    # 'ce' transport: who moves the blocks - "kernel" (evx_peer_scatter, a few CTAs per peer on
    # a side stream) or "dma" (one cudaMemcpyAsync per peer and chunk; fine for 2 GPUs, but the
    # per-copy latency of the DMA engines adds up with 7 peers)
    copier = "dma"
    scatter_ctas = 8
    # the copies of the LAST chunk of a pipelined stage overlap with nothing: move them with an SM
    # copy kernel that fills the GPU (NVLink-bound, ~700 GB/s aggregate) instead of the DMA engines
    # (measured ~420 GB/s aggregate over 7 peers for contiguous regions, ~215 GB/s for the pitched
    # regions after the x pass).  0 = DMA for every chunk.
    last_chunk_ctas = int(os.environ.get("EVX_CE_LAST_CTAS", "24"))
    # the DMA copies of a chunk as ONE unordered batch (cudaMemcpyBatchAsync): 1 = forward and x pass,
    # 2 = forward only
    batch_copies = int(os.environ.get("EVX_CE_BATCH", "0"))
    # hybrid transport: number of peers served by TMA stores from inside the y pass / x pass
    direct_peers = int(os.environ.get("EVX_CE_DIRECT", "0"))
    direct_peers_mid = int(os.environ.get("EVX_CE_DIRECT_MID", os.environ.get("EVX_CE_DIRECT", "0")))

    def _mark(self, name, stream=None):
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream if stream is not None else torch.cuda.current_stream(self.device))
            self.trace.append((name, e))

    # copy-engine transport: kernels write local block buffers, DMA engines move the blocks ---
    def halo_stream(self):
        if not hasattr(self, "_halo"):
            self._halo = torch.cuda.Stream(device=self.device)
        return self._halo

    def _ce_streams(self):
        if not hasattr(self, "_comp"):
            self._comp = torch.cuda.Stream(device=self.device)
            self._link = torch.cuda.Stream(device=self.device)
        if not hasattr(self, "_copy_streams"):
            self._copy_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.slab.world)]
        return self._comp, self._copy_streams

    @staticmethod
    def _split(n, chunks, env, default=None):
        """Chunk bounds of a range of n items: `chunks` equal parts, or the fractions in the
        environment variable `env` (comma-separated, any positive weights)."""
        spec = os.environ.get(env, default)
        if spec:
            w = [float(v) for v in spec.split(",") if v.strip()]
            if len(w) >= 1 and all(v > 0 for v in w) and n >= 8 * len(w):
                acc, tot, b = 0.0, sum(w), [0]
                for v in w[:-1]:
                    acc += v
                    b.append(min(n, max(b[-1] + 1, round(n * acc / tot))))
                return b + [n]
        return [round(i * n / chunks) for i in range(chunks + 1)]

    def forward_ce(self, u, rhs, eps, D, bc, halo_lo, halo_hi, chunks, halo_event=None):
        """rhs -> z pass -> y pass into the local block buffer A over `chunks` slices of the
        local x range (the block that stays on this GPU goes straight into the local B); as
        soon as a slice is done its rows of every other block are copied by the DMA engines
        into block `rank` of the owning rank's buffer B (one copy stream per peer) while the
        kernels of the next slice run.  The slices that need no halo planes go first, so that the
        halo exchange (`halo_event`: recorded behind it on its own stream) hides under their
        kernels.  Returns the local B."""
        W, me, nxl, nyl, P = self.slab.world, self.slab.rank, self.slab.nxl, self.slab.nyl, self.plan.pitch
        blk = nxl * nyl * P * 8                      # bytes per block
        row = nyl * P * 8                            # bytes per local x plane inside a block
        # 2 GPUs (kernel-bound): thin edge slices - they run last (they wait for the halos) and the
        # copies of the very last slice overlap with nothing (4 slices: 1.98 -> 1.92 ms/step against
        # equal slices).  8 GPUs (copy-bound): equal slices (3.00 against 3.19 ms/step).
        bounds = self._split(nxl, chunks, "EVX_CE_FWD_SPLIT", "0.12,0.38,0.38,0.12" if chunks == 4 and W == 2 else None)
        chunks = len(bounds) - 1
        order = [i for i in range(chunks) if 0 < i < chunks - 1] + sorted({0, chunks - 1})
        if halo_event is None or os.environ.get("EVX_CE_HALO_OVERLAP", "1") == "0":
            order = list(range(chunks))
        main = torch.cuda.current_stream(self.device)
        comp, copies = self._ce_streams()
        comp.wait_stream(main)
        halo_waited = halo_event is None
        a_ptr, b_ptrs = self.buf_a.data_ptr(), self.peers.peer_ptrs[1]
        stagger = [(me + 1 + k) % W for k in range(W - 1)]        # staggered across ranks
        # hybrid transport: the blocks of the first `direct_peers` ranks of that order leave the y
        # pass as TMA stores into the peer's buffer (no send-buffer round trip, no DMA copy), the
        # others go through the local block buffer and the copy engines as before
        direct = stagger[:max(0, min(self.direct_peers, W - 1))]
        table = [b_ptrs[j] if j in direct else (self.buf_b.data_ptr() if j == me else a_ptr + (j - me) * blk)
                 for j in range(W)]
        if direct:
            self.plan.set_p2p_ctas(0)
        for n_done, i in enumerate(order):
            x0, x1 = bounds[i], bounds[i + 1]
            with torch.cuda.stream(comp):
                if not halo_waited and (x0 == 0 or x1 == nxl):
                    comp.wait_event(halo_event)
                    halo_waited = True
                lo = halo_lo if x0 == 0 else u[x0 - 2:x0]
                hi = halo_hi if x1 == nxl else u[x1:x1 + 2]
                _native.ch_rhs(u[x0:x1], rhs[x0:x1], self.spacing, eps, D, bc, halo_lo=lo, halo_hi=hi)
                self._mark(f"fwd{i} rhs", comp)
                if direct:
                    self.plan.forward_chunk_p2p(rhs, self.spec, table, x0, x1 - x0, parts=3)
                else:
                    self.plan.forward_chunk(rhs, self.spec, self.buf_a, x0, x1 - x0, self_block=self.buf_b)
                self._mark(f"fwd{i} z+y", comp)
                done = torch.cuda.Event()
                done.record(comp)
            peers_ = [j for j in stagger if j not in direct]
            last = n_done == chunks - 1 and self.last_chunk_ctas > 0 and W > 1
            if not peers_:
                continue
            if self.copier == "kernel" or last:
                cs = comp if last else copies[0]
                cs.wait_event(done)
                _native.peer_scatter([a_ptr + j * blk + x0 * row for j in peers_],
                                     [b_ptrs[j] + me * blk + x0 * row for j in peers_],
                                     (x1 - x0) * row, 1, (x1 - x0) * row, (x1 - x0) * row,
                                     self.last_chunk_ctas if last else self.scatter_ctas, cs)
                self._mark(f"fwd{i} scatter", cs)
                continue
            if self.batch_copies:
                cs = copies[0]
                cs.wait_event(done)
                _native.copy_batch_async([b_ptrs[j] + me * blk + x0 * row for j in peers_], (x1 - x0) * row,
                                         [a_ptr + j * blk + x0 * row for j in peers_], (x1 - x0) * row,
                                         (x1 - x0) * row, 1, cs)
                self._mark(f"fwd{i} batch", cs)
                continue
            for j in peers_:
                cs = copies[j]
                cs.wait_event(done)
                _native.copy_async(b_ptrs[j] + me * blk + x0 * row, a_ptr + j * blk + x0 * row,
                                   (x1 - x0) * row, cs)
                self._mark(f"fwd{i} copy->{j}", cs)
        if not halo_waited:
            comp.wait_event(halo_event)
        main.wait_stream(comp)
        for cs in copies:
            main.wait_stream(cs)
        self._mark("fwd joined")
        self.peers.barrier(1)
        self._mark("fwd barrier")
        return self.buf_b

    def middle_ce(self, dt, coef, power, chunks):
        """x pass in place on the local B over `chunks` slices of the local y-pencil rows; each
        finished slice is copied (2-D region per block) into block `rank` of the owning
        rank's buffer A.  Returns the local A."""
        W, me, nxl, nyl, P = self.slab.world, self.slab.rank, self.slab.nxl, self.slab.nyl, self.plan.pitch
        blk = nxl * nyl * P * 8
        pitch = nyl * P * 8
        chunks = max(1, min(chunks, nyl))
        bounds = self._split(nyl, chunks, "EVX_CE_MID_SPLIT")
        chunks = len(bounds) - 1
        main = torch.cuda.current_stream(self.device)
        comp, copies = self._ce_streams()
        comp.wait_stream(main)
        b_ptr, a_ptrs = self.buf_b.data_ptr(), self.peers.peer_ptrs[0]
        stagger = [(me + 1 + k) % W for k in range(W - 1)]
        direct = stagger[:max(0, min(self.direct_peers_mid, W - 1))]      # see forward_ce
        table = [a_ptrs[j] if j in direct else (self.buf_a.data_ptr() if j == me else b_ptr + (j - me) * blk)
                 for j in range(W)]
        if direct:
            self.plan.set_p2p_ctas(0)
        for i in range(chunks):
            y0, y1 = bounds[i], bounds[i + 1]
            with torch.cuda.stream(comp):
                if direct:
                    self.plan.middle_chunk_p2p(self.buf_b, table, y0, y1 - y0, self.spacing, dt, coef, power)
                else:
                    self.plan.middle_chunk(self.buf_b, y0, y1 - y0, self.spacing, dt, coef, power,
                                           self_block=self.buf_a)
                self._mark(f"mid{i} x", comp)
                done = torch.cuda.Event()
                done.record(comp)
            peers_ = [j for j in stagger if j not in direct]
            last = i == chunks - 1 and self.last_chunk_ctas > 0 and W > 1
            if not peers_:
                continue
            if self.copier == "kernel" or last:
                cs = comp if last else copies[0]
                cs.wait_event(done)
                _native.peer_scatter([b_ptr + j * blk + y0 * P * 8 for j in peers_],
                                     [a_ptrs[j] + me * blk + y0 * P * 8 for j in peers_],
                                     (y1 - y0) * P * 8, nxl, pitch, pitch,
                                     self.last_chunk_ctas if last else self.scatter_ctas, cs)
                self._mark(f"mid{i} scatter", cs)
                continue
            if self.batch_copies == 1:
                cs = copies[0]
                cs.wait_event(done)
                if y1 - y0 == nyl:
                    _native.copy_batch_async([a_ptrs[j] + me * blk for j in peers_], blk,
                                             [b_ptr + j * blk for j in peers_], blk, blk, 1, cs)
                else:
                    _native.copy_batch_async([a_ptrs[j] + me * blk + y0 * P * 8 for j in peers_], pitch,
                                             [b_ptr + j * blk + y0 * P * 8 for j in peers_], pitch,
                                             (y1 - y0) * P * 8, nxl, cs)
                self._mark(f"mid{i} batch", cs)
                continue
            for j in peers_:
                cs = copies[j]
                cs.wait_event(done)
                if y1 - y0 == nyl:       # whole block: one contiguous copy (fastest DMA path)
                    _native.copy_async(a_ptrs[j] + me * blk, b_ptr + j * blk, blk, cs)
                else:
                    _native.copy2d_async(a_ptrs[j] + me * blk + y0 * P * 8, pitch,
                                         b_ptr + j * blk + y0 * P * 8, pitch, (y1 - y0) * P * 8, nxl, cs)
                self._mark(f"mid{i} copy->{j}", cs)
        main.wait_stream(comp)
        for cs in copies:
            main.wait_stream(cs)
        self._mark("mid joined")
        self.peers.barrier(0)
        self._mark("mid barrier")
        return self.buf_a

    def spectral_middle_p2p(self, dt, coef, power):
        """x pass on the local B; chunks land in every peer's buffer A.  Returns the local A."""
        self.plan.middle_p2p(self.buf_b, self.peers.peer_ptrs[0], self.spacing, dt, coef, power)
        self.peers.barrier(0)
        return self.buf_a

    def spectral_backward(self, buf, u, out):
        self.plan.backward(buf, self.spec, u, out)

    def exchange_buffers(self):
        return self.buf_a, self.buf_b


class DistributedCahnHilliardIMEX:
    """CahnHilliard + PseudoSpectralIMEX (fully periodic) on an x-slab decomposition.
    `step(u_local) -> u_local_new`; same arithmetic as the single-GPU step."""

    def __init__(self, global_shape, spacing, dt, eps=3.0, D=1.0, A=0.25, group=None,
                 device=None, ops=None, transport="ce", overlap_chunks=4, p2p_ctas=148,
                 copier=None, scatter_ctas=None, mid_chunks=None, hom_fn=None):
        # hom_fn: user potential, [1, n, ny, nz] raw c -> mu_hom(clip(c)) (CahnHilliard.hom_field);
        # None = the default double well, evaluated inside the rhs kernel
        self.hom_fn = hom_fn
        self.comm = Comm(group)
        self.slab = Slab(tuple(global_shape), self.comm.world, self.comm.rank)
        self.spacing, self.dt, self.eps, self.D, self.A = tuple(spacing), dt, eps, D, A
        self.bc = normalize_bc(("periodic",) * 3)
        self.ops = ops if ops is not None else CudaOps(self.slab, spacing, device or "cuda",
                                                       transport=transport, group=group)
        # x-pass pipeline depth of the 'ce' transport (measured on 8 GPUs, 512^3 per GPU:
        # 4 chunks 3.39 ms/step, 2 chunks 3.40, unchunked 3.94)
        # (2 GPUs, round 2: 8 chunks 1.79 ms/step, 6 chunks 1.80, 4 chunks 1.82, 3 chunks 1.85; 8 GPUs: 4 chunks)
        self.mid_chunks = mid_chunks if mid_chunks else (8 if self.comm.world == 2 and overlap_chunks == 4
                                                         else overlap_chunks)
        if (self.comm.world >= 8 and ops is None and "EVX_CE_DIRECT" not in os.environ
                and tuple(global_shape[:2]) == (1024, 1024)):      # measured with the TMA-store passes only
            # 8 GPUs (copy-bound): two of the seven blocks leave the passes as TMA stores over NVLink
            # instead of through the copy engines (3.03 -> 2.93 ms/step; 1: 2.99, 3: 3.01)
            self.ops.direct_peers = self.ops.direct_peers_mid = 2
        if self.comm.world == 2 and "EVX_CE_LAST_CTAS" not in os.environ and ops is None:
            # one peer: the last slice's blocks move fastest with the whole GPU copying (1.81 -> 1.77 ms)
            self.ops.last_chunk_ctas = 148
        if copier is not None:
            self.ops.copier = copier
        if scatter_ctas:
            self.ops.scatter_ctas = int(scatter_ctas)
        self.rhs = self.ops.new_field()
        # forward pipeline depth: x chunks of the local slab issued on two streams (p2p only)
        self.overlap_chunks = overlap_chunks if self.slab.nxl >= 8 * max(overlap_chunks, 1) else 1
        if ops is None and self.ops.transport == "p2p" and self.overlap_chunks > 1 and p2p_ctas:
            self.ops.plan.set_p2p_ctas(p2p_ctas)

    def step_profiled(self, u_local):
        """One step with CUDA events around every stage (diagnostics; p2p transport,
        un-pipelined).  Returns (u_new, {stage: ms})."""
        ops, comm = self.ops, self.comm
        ev = []

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append((name, e))

        u_local = u_local.contiguous()
        mark("start")
        halo_lo, halo_hi = comm.exchange_halos(u_local, 2, periodic=True)
        mark("halo_exchange")
        ops.ch_rhs(u_local, self.rhs, self.eps, self.D, self.bc, halo_lo, halo_hi)
        mark("ch_rhs")
        coef = 2.0 * self.eps * self.D * self.A
        ops.plan.forward_p2p(self.rhs, ops.spec, ops.peers.peer_ptrs[1])
        mark("z_fwd+y_fwd(p2p stores)")
        ops.peers.barrier(1)
        mark("barrier")
        ops.plan.middle_p2p(ops.buf_b, ops.peers.peer_ptrs[0], self.spacing, self.dt, coef, 2)
        mark("x_mid(p2p stores)")
        ops.peers.barrier(0)
        mark("barrier2")
        out = ops.new_field()
        ops.spectral_backward(ops.buf_a, u_local, out)
        mark("y_inv+z_inv")
        torch.cuda.synchronize()
        times = {ev[i][0]: ev[i - 1][1].elapsed_time(ev[i][1]) for i in range(1, len(ev))}
        return out, times

    def _step_user_potential(self, u_local):
        """Custom mu_hom: the potential is a caller-evaluated field, which the fused kernel
        accepts only without x halos - so the rhs runs on the slab extended by its two halo
        planes per side (periodic in x inside the kernel: the wrap only reaches the four outer
        planes, which are dropped) and the update goes through the un-pipelined transposes."""
        ops, comm = self.ops, self.comm
        halo_lo, halo_hi = comm.exchange_halos(u_local, 2, periodic=True)
        if halo_lo is None:            # one rank: the kernel wraps by index arithmetic
            ops.ch_rhs_hom(u_local, self.rhs, self.eps, self.D, self.bc,
                           self.hom_fn(u_local[None])[0].contiguous())
        else:
            ext = torch.cat([halo_lo, u_local, halo_hi], 0).contiguous()
            hom = self.hom_fn(ext[None])[0].contiguous()
            rhs_ext = torch.empty_like(ext)
            ops.ch_rhs_hom(ext, rhs_ext, self.eps, self.D, self.bc, hom)
            self.rhs.copy_(rhs_ext[2:-2])
        coef = 2.0 * self.eps * self.D * self.A
        out = ops.new_field()
        if getattr(ops, "transport", "nccl") in ("p2p", "ce") and ops.peers is not None:
            ops.spectral_forward_p2p(self.rhs)
            a = ops.spectral_middle_p2p(self.dt, coef, 2)
            ops.spectral_backward(a, u_local, out)
            return out
        a, b = ops.exchange_buffers()
        send = ops.spectral_forward(self.rhs)
        comm.all_to_all_blocks(send, b)
        ops.spectral_middle(b, self.dt, coef, 2)
        comm.all_to_all_blocks(b, a)
        ops.spectral_backward(a, u_local, out)
        return out

    def step(self, u_local):
        ops, comm = self.ops, self.comm
        u_local = u_local.contiguous()
        if self.hom_fn is not None:
            return self._step_user_potential(u_local)
        transport = getattr(ops, "transport", "nccl")
        halo_event = None
        if transport == "ce":
            # halo planes travel on their own stream; the chunks that need them run last
            main = torch.cuda.current_stream(u_local.device)
            hs = ops.halo_stream()
            hs.wait_stream(main)
            with torch.cuda.stream(hs):
                halo_lo, halo_hi = ops.halos.exchange(u_local)
                ops._mark("halo done", hs)
                halo_event = torch.cuda.Event()
                halo_event.record(hs)
        elif transport == "p2p" and getattr(ops, "halos", None) is not None:
            halo_lo, halo_hi = ops.halos.exchange(u_local)
        else:
            halo_lo, halo_hi = comm.exchange_halos(u_local, 2, periodic=True)
        if transport == "ce":
            ops.forward_ce(u_local, self.rhs, self.eps, self.D, self.bc, halo_lo, halo_hi,
                           max(self.overlap_chunks, 1), halo_event=halo_event)
            a = ops.middle_ce(self.dt, 2.0 * self.eps * self.D * self.A, 2, max(self.mid_chunks, 1))
            out = ops.new_field()
            ops.spectral_backward(a, u_local, out)
            ops._mark("bwd y+z")
            return out
        pipelined = transport == "p2p" and self.overlap_chunks > 1
        if not pipelined:
            ops.ch_rhs(u_local, self.rhs, self.eps, self.D, self.bc, halo_lo, halo_hi)
        coef = 2.0 * self.eps * self.D * self.A
        if getattr(ops, "transport", "nccl") == "p2p":
            if self.overlap_chunks > 1:
                # the rhs was NOT computed above in this mode (see below): pipeline
                # rhs -> z pass -> y pass(+NVLink stores) over x chunks on two streams
                ops.forward_pipelined(u_local, self.rhs, self.eps, self.D, self.bc, halo_lo,
                                      halo_hi, self.overlap_chunks)
            else:
                ops.spectral_forward_p2p(self.rhs)
            a = ops.spectral_middle_p2p(self.dt, coef, 2)
            out = ops.new_field()
            ops.spectral_backward(a, u_local, out)
            return out
        a, b = ops.exchange_buffers()
        send = ops.spectral_forward(self.rhs)            # fills `a`
        comm.all_to_all_blocks(send, b)
        ops.spectral_middle(b, self.dt, 2.0 * self.eps * self.D * self.A, 2)
        comm.all_to_all_blocks(b, a)
        out = ops.new_field()
        ops.spectral_backward(a, u_local, out)
        return out

    def total_mass(self, u_local):
        s = u_local.double().sum().reshape(1)
        return float(self.comm.all_reduce_sum(s).item())

    # ---- adjoint of one step (last row: same partitioning, parameter-gradient
    # partial sums all-reduced once) ---------------------------------------------------------
    ADJ_HALO = 4      # planes per side of the extended slab the adjoint stencil runs on

    def spectral_filter(self, r_local):
        """G r = F^-1[ dt / (1 + dt s) F r ] of a slab-decomposed field (self-adjoint: the same
        call serves the forward update and w = G lam+ of the backward pass)."""
        ops, comm = self.ops, self.comm
        coef = 2.0 * self.eps * self.D * self.A
        out = ops.new_field()
        r_local = r_local.contiguous()
        if getattr(ops, "transport", "nccl") in ("p2p", "ce") and ops.peers is not None:
            ops.spectral_forward_p2p(r_local)
            a = ops.spectral_middle_p2p(self.dt, coef, 2)
            ops.spectral_backward(a, None, out)
            return out
        a, b = ops.exchange_buffers()
        send = ops.spectral_forward(r_local)
        comm.all_to_all_blocks(send, b)
        ops.spectral_middle(b, self.dt, coef, 2)
        comm.all_to_all_blocks(b, a)
        ops.spectral_backward(a, None, out)
        return out

    def _extend(self, f_local, width):
        lo, hi = self.comm.exchange_halos(f_local, width, periodic=True)
        if lo is None:                       # one rank: the kernels wrap by index arithmetic
            return f_local, 0
        return torch.cat([lo, f_local, hi], 0).contiguous(), width

    def step_vjp(self, u_local, u_plus_local, lam_plus_local):
        """Backward pass of `step` on this rank's slab: given lam+ = dL/du+ returns
        (dL/du, dL/dD, dL/deps) - the field gradient of the slab, the two parameter gradients
        ALREADY summed over all ranks (one all-reduce of a 2-vector).  Same formulas as the
        single-GPU adjoint (autograd._CHImexStepFn):
            w = G lam+ (distributed filter), lam = lam+ + dR/du^T w (stencil kernels on the slab
            extended by ADJ_HALO planes of u and w per side),
            dL/dD = <w, u+ - u> / (dt D),  dL/deps = <z, dmu/deps> + <w/dt - lam+, u+ - u> / eps."""
        if self.hom_fn is not None:
            raise NotImplementedError("the hand-written adjoint supports the default mu_hom only")
        u_local, lam_plus_local = u_local.contiguous(), lam_plus_local.contiguous()
        if self.slab.world > 1 and self.slab.nxl < self.ADJ_HALO:
            raise ValueError("slab thinner than the adjoint halo")
        w = self.spectral_filter(lam_plus_local)
        u_ext, h = self._extend(u_local, self.ADJ_HALO)
        w_ext, _ = self._extend(w, self.ADJ_HALO)
        n = u_local.shape[0]
        if h:
            lam_in = torch.zeros_like(u_ext)
            lam_in[h:h + n] = lam_plus_local
        else:
            lam_in = lam_plus_local
        lam_ext, deps = self.ops.ch_rhs_vjp_ext(u_ext, w_ext, lam_in, self.eps, self.D, h, h + n)
        lam = lam_ext[h:h + n].contiguous() if h else lam_ext
        du = (u_plus_local - u_local).double()
        wd = w.double()
        g_D = (wd * du).sum() / (self.dt * self.D)
        g_eps = deps.double() + ((wd / self.dt - lam_plus_local.double()) * du).sum() / self.eps
        g = self.comm.all_reduce_sum(torch.stack([g_D, g_eps]))
        return lam, g[0], g[1]

    def step_autograd(self, u_local, D=None, eps=None):
        """Differentiable `step`: u_local (and optionally 0-dim tensors D, eps that carry
        requires_grad) -> u+ of the slab, with `step_vjp` as the backward pass.  Every rank
        must call it (and later backward) collectively."""
        like = u_local
        D_t = D if isinstance(D, torch.Tensor) else torch.tensor(float(self.D), dtype=torch.float64, device=like.device)
        e_t = eps if isinstance(eps, torch.Tensor) else torch.tensor(float(self.eps), dtype=torch.float64, device=like.device)
        return _DistCHImexStepFn.apply(u_local.contiguous(), D_t, e_t, self)


class _DistCHImexStepFn(torch.autograd.Function):
    """One node per distributed step; saves (u, u+) of the slab only."""

    @staticmethod
    def forward(ctx, u, D, eps, stepper):
        stepper.D, stepper.eps = float(D), float(eps)
        out = stepper.step(u.detach())
        ctx.save_for_backward(u, out, D, eps)
        ctx.stepper = stepper
        return out

    @staticmethod
    def backward(ctx, lam_plus):
        u, u_plus, D, eps = ctx.saved_tensors
        st = ctx.stepper
        st.D, st.eps = float(D), float(eps)
        lam, g_D, g_eps = st.step_vjp(u.detach(), u_plus, lam_plus)
        return lam, g_D.to(D.dtype), g_eps.to(eps.dtype), None


class DistributedAllenCahnEuler:
    """TwoPhaseAllenCahn + ForwardEuler on an x-slab decomposition (1-plane halos)."""

    def __init__(self, global_shape, spacing, dt, eps=2.0, gab=1.0, M=1.0, force=0.0,
                 curvature=0.01, bc=("neumann",) * 3, group=None, device=None, ops=None):
        self.comm = Comm(group)
        self.slab = Slab(tuple(global_shape), self.comm.world, self.comm.rank)
        self.dt = dt
        self.params = dict(eps=eps, gab=gab, M=M, force=force, curvature=curvature)
        self.bc = normalize_bc(bc)
        self.ops = ops if ops is not None else CudaOps(self.slab, spacing, device or "cuda",
                                                       spectral=False)
        # halos as DMA writes into the neighbours' symmetric-memory slots where that exists
        # (CUDA ranks of one NVSwitch box); otherwise torch.distributed send / recv
        self.halos = None
        if ops is None and self.slab.world > 1:
            try:
                _, ny, nz = self.slab.global_shape
                self.halos = PeerHalos(1, ny, nz, torch.device(device or "cuda"), group, self.slab.world,
                                       self.slab.rank)
            except Exception as exc:
                import warnings
                warnings.warn(f"torch symmetric memory is not available ({exc}); Allen-Cahn halos use "
                              "send / recv")

    def step(self, phi_local):
        phi_local = phi_local.contiguous()
        periodic = self.bc[0][0] == "periodic"
        if self.halos is not None:
            halo_lo, halo_hi = self.halos.exchange(phi_local, periodic=periodic)
        else:
            halo_lo, halo_hi = self.comm.exchange_halos(phi_local, 1, periodic=periodic)
        out = torch.empty_like(phi_local)
        self.ops.ac_stage(phi_local, out, self.params, self.bc, self.dt, halo_lo, halo_hi)
        return out
