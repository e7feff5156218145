#!/usr/bin/env python
"""Benchmark of the evoxels hot path: one semi-implicit Cahn-Hilliard step per "step".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl reference]

N = 1 (default) workload: BASELINE.json configs[1] - Cahn-Hilliard IMEX, 512^3, fp32,
periodic, h = 1, dt = 0.1, eps = 3, D = 1, A = 0.25, c0 = 0.5 + 0.1 U[0,1) seed 0.
Prints ONE JSON line (see the contract in the task description):
  value      voxel-updates/s with the state resident in HBM (CUDA-event time, max over ranks)
  e2e        same metric through the public step call with PINNED HOST buffers: H2D of the
             field, one step, D2H of the result, every step
  roofline   algorithmic bytes / event time of the dominant own kernel, + whole-step figure
  cpu_baseline  the oracle port (reference algorithm, torch CPU) timed on this box's cores
`--impl reference` times that CPU port alone (the reference is pure Python/torch; its CPU
path *is* this arithmetic - see oracle/evx_oracle.py) and prints the same line shape.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voxel-updates/sec per CH step"
UNIT = "voxel-updates/s"
CH = dict(eps=3.0, D=1.0, A=0.25, dt=0.1)
B_ALG_STEP = 60.0       # algorithmic bytes / voxel / step
B_ALG_RHS = 8.0         # fused rhs kernel: read c, write rhs
# dram__bytes_read.sum + dram__bytes_write.sum per launch at 512^3 from the committed
# `ncu --set full` capture (profiles/r01_ncu_full_summary_final.txt); bytes
NCU_TRAFFIC_512 = {
    "ch_rhs_kernel": 0.5411e9 + 0.5019e9,
    "fft_z_forward (ZPass fwd)": 0.5446e9 + 0.5025e9,
    "fft_y_forward (StridedPipe FWD)": 0.5540e9 + 0.4933e9,
    "fft_x_fwd*filter*inv (StridedPipe XMID)": 0.5537e9 + 0.4871e9,
    "fft_y_inverse (StridedPipe INV)": 0.5540e9 + 0.4933e9,
    "fft_z_inverse+u (ZPass inv)": 1.091e9 + 0.5084e9,
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.gpu)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
        try:
            sm, mx, reasons = [], [], set()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
            if sm:
                out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx),
                           reasons=sorted(reasons), samples=len(sm))
        except Exception:
            pass
        return out



def e2e_pipelined(step_fn, h_in, h_outs, dev, nsteps):
    """End-to-end loop: every step copies its input field from pinned host memory to the device,
    runs one step and copies the result back to pinned host memory.  The three legs of
    consecutive (independent) requests overlap on three streams - H2D of request i+1, the step
    of request i and D2H of request i-1 - because PCIe is full duplex and the copy engines are
    idle while the SMs work.  Returns the CUDA-event time of the whole loop in ms."""
    import torch
    s_in, s_run, s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
    d_in = [torch.empty(h_in.shape, dtype=h_in.dtype, device=dev) for _ in range(2)]
    ran = [None, None]        # event: step that consumed d_in[k] has finished
    copied = [None, None]     # event: h_outs[k] has been filled
    main = torch.cuda.current_stream(dev)
    for s in (s_in, s_run, s_out):
        s.wait_stream(main)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(s_in)
    for i in range(nsteps):
        k = i % 2
        with torch.cuda.stream(s_in):
            if ran[k] is not None:
                s_in.wait_event(ran[k])
            d_in[k].copy_(h_in, non_blocking=True)
            arrived = torch.cuda.Event()
            arrived.record(s_in)
        with torch.cuda.stream(s_run):
            s_run.wait_event(arrived)
            out = step_fn(d_in[k])
            ran[k] = torch.cuda.Event()
            ran[k].record(s_run)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ran[k])
            if copied[k] is not None:
                s_out.wait_event(copied[k])
            h_outs[k].copy_(out, non_blocking=True)
            out.record_stream(s_out)
            copied[k] = torch.cuda.Event()
            copied[k].record(s_out)
    main.wait_stream(s_in)
    main.wait_stream(s_run)
    main.wait_stream(s_out)
    t1.record(main)
    torch.cuda.synchronize(dev)
    return t0.elapsed_time(t1)

def time_cpu_port(size, steps, warmup):
    """Oracle port of the reference step on the host cores. Returns (vox/s, s/step, threads)."""
    import torch
    from oracle import evx_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    shape = (size, size, size)
    orc = O.CHOracle(shape, (1.0, 1.0, 1.0), CH["dt"], CH["eps"], CH["D"], CH["A"])
    u = O.noise_field(shape, seed=0)
    for _ in range(warmup):
        u = orc.step(u)
    t0 = time.perf_counter()
    for _ in range(steps):
        u = orc.step(u)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return size ** 3 / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: the largest cube whose (K+W) steps fit ~150 s on these cores
    _, t128, _ = time_cpu_port(128, 1, 1)
    total = args.steps + args.warmup
    size = 128
    for cand, factor in ((512, 64 * 3.0), (256, 8 * 2.5)):   # voxel ratio x cache penalty
        if t128 * factor * total <= 150.0:
            size = cand
            break
    vps, spstep, threads = time_cpu_port(size, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": spstep * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"CH IMEX {size}^3 fp32 periodic dt=0.1 (CPU, bounded sample of the 512^3 workload)",
                   **CH},
        "cpu_baseline": {"value": vps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of {size}^3 after {args.warmup} warm-up, torch CPU eager"},
        "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import evoxels_b200 as evo
    from evoxels_b200 import _native
    from evoxels_b200.problem_definition import CahnHilliard
    from evoxels_b200.timesteppers import PseudoSpectralIMEX
    from evoxels_b200.voxelgrid import VoxelGridTorch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.size
    if world > 1:
        return run_ours_distributed(args, world, rank, local, dev)
    shape = (n, n, n)
    nvox = n ** 3
    vf = evo.VoxelFields(shape, tuple(float(s) for s in shape))
    vg = VoxelGridTorch(vf.grid_info(), device=str(dev))
    prob = CahnHilliard(vg, eps=CH["eps"], D=CH["D"], A=CH["A"])
    ts = PseudoSpectralIMEX(prob, CH["dt"], fft_backend=args.fft)
    gen = torch.Generator(device=dev).manual_seed(rank)
    u0 = 0.5 + 0.1 * torch.rand((1,) + shape, device=dev, generator=gen)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput -----------------------------------------------------
    u = u0
    for _ in range(max(args.warmup, 3)):
        u = ts.step(0.0, u)
    plan = next(iter(ts._plans.values()))
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        u = ts.step(0.0, u)
    e1.record()
    barrier()
    launches = _native.launch_count() - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else {}
    ms_step = ms / args.steps
    value = world * nvox * args.steps / (ms * 1e-3)
    mass_drift = abs(float(u.double().mean()) - float(u0.double().mean()))

    # ---- end to end: pinned host -> device -> step -> pinned host, every step ------------
    e2e_steps = max(6, min(args.steps, 20))
    h_in = torch.empty((1,) + shape, dtype=torch.float32, device="cpu", pin_memory=True)
    h_in.copy_(u0)
    h_outs = [torch.empty((1,) + shape, dtype=torch.float32, device="cpu", pin_memory=True)
              for _ in range(2)]
    e2e_pipelined(lambda d: ts.step(0.0, d), h_in, h_outs, dev, 3)       # warm-up
    barrier()
    ms_e2e = e2e_pipelined(lambda d: ts.step(0.0, d), h_in, h_outs, dev, e2e_steps)
    barrier()
    e2e_check = float((h_outs[(e2e_steps - 1) % 2].to(dev) - ts.step(0.0, u0)).abs().max())
    e2e_value = world * nvox * e2e_steps / (ms_e2e * 1e-3)
    field_bytes = nvox * 4

    # ---- per-kernel timing of the own kernels (roofline) -----------------------------------
    peak, peak_src = measured_peaks()
    rhs_buf = torch.empty_like(u0)
    out_buf = torch.empty_like(u0)
    reps = max(5, min(args.steps, 20))

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    per = (("periodic", None),) * 3
    coef = 2 * CH["eps"] * CH["D"] * CH["A"]
    half = (n // 2 + 1) / (n // 2)          # half spectrum: (nz/2+1)/(nz/2) x 4 B per voxel
    kernels = {}                            # name -> (ms, algorithmic bytes per voxel)
    kernels["ch_rhs_kernel"] = (
        timed(lambda: _native.ch_rhs(u0[0], rhs_buf[0], vg.spacing, CH["eps"], CH["D"], per)), 8.0)
    if plan.backend_name == "native":
        names = ["fft_z_forward (ZPass fwd)", "fft_y_forward (StridedPipe FWD)",
                 "fft_x_fwd*filter*inv (StridedPipe XMID)", "fft_y_inverse (StridedPipe INV)",
                 "fft_z_inverse+u (ZPass inv)"]
        bpv = [4 + 4 * half, 8 * half, 8 * half, 8 * half, 4 * half + 8]
        for which, (nm, bb) in enumerate(zip(names, bpv)):
            kernels[nm] = (timed(lambda w=which: plan.native_pass(
                w, u0[0], rhs_buf[0], out_buf[0], vg.spacing, CH["dt"], coef, 2)), bb)
    else:
        kernels["spectral_apply (cuFFT + filter + add kernels)"] = (
            timed(lambda: plan.apply(u0[0], rhs_buf[0], out_buf[0], vg.spacing, CH["dt"], coef, 2)), 52.0)
    dom = max(kernels, key=lambda k: kernels[k][0])
    dom_ms, dom_bpv = kernels[dom]
    dom_gbs = dom_bpv * nvox / (dom_ms * 1e-3) / 1e9
    gbs_step = B_ALG_STEP * nvox / (ms_step * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak, "unit": "GB/s",
        "frac": dom_gbs / peak, "traffic": NCU_TRAFFIC_512.get(dom) if n == 512 else None,
        "traffic_source": "profiles/r01_ncu_full_summary_final.txt (ncu --set full, per launch)",
        "peak_source": peak_src, "ms": dom_ms,
        "bytes_per_voxel": dom_bpv,
        "step": {"bytes_per_voxel": B_ALG_STEP, "achieved": gbs_step, "frac": gbs_step / peak,
                 "frac_of_8TBs": gbs_step / 8000.0, "ms": ms_step},
        "kernels": {k: {"ms": v[0], "bytes_per_voxel": v[1],
                        "achieved_GBs": v[1] * nvox / (v[0] * 1e-3) / 1e9,
                        "frac": v[1] * nvox / (v[0] * 1e-3) / 1e9 / peak} for k, v in kernels.items()},
    }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample) -------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            cs = 512 if n >= 512 else n
            vps, spstep, threads = time_cpu_port(cs, 2, 1)
            cpu = {"value": vps, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"2 steps of {cs}^3 after 1 warm-up (oracle port of the reference, torch CPU eager)",
                   "s_per_step": spstep}
        except Exception as exc:  # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"failed: {exc}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CH IMEX {n}^3 fp32 periodic dt=0.1 per GPU", **CH,
                   "fft_backend": plan.backend_name,
                   "parallelism": "single GPU" if world == 1 else f"{world} independent replicas",
                   "l2": "field (%.0f MB) larger than L2 (126 MB), no flush needed" % (field_bytes / 1e6)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": field_bytes,
                "d2h_bytes_per_step": field_bytes, "ms_per_step": ms_e2e / e2e_steps,
                "steps": e2e_steps,
                "max_abs_diff_vs_device_step": e2e_check,
                "note": "pinned host field -> H2D -> PseudoSpectralIMEX.step -> D2H, every step; "
                        "the three legs of consecutive steps overlap on three streams"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "mass_drift": mass_drift,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def weak_scaling_shape(n, world):
    """Global grid with n^3 voxels per GPU, all extents powers of two <= 2048:
    1: n^3, 2: (2n,n,n), 4: (2n,2n,n), 8: (2n,2n,2n)."""
    f = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    return tuple(n * k for k in f)


def run_ours_distributed(args, world, rank, local, dev):
    """N > 1: ONE global Cahn-Hilliard problem, x-slab decomposed over the ranks (weak
    scaling: args.size^3 voxels per GPU): halo planes and slab<->pencil transposes over
    NVLink (see --transport)."""
    import torch
    import torch.distributed as dist
    from evoxels_b200 import _native
    from evoxels_b200.distributed import DistributedCahnHilliardIMEX

    shape = weak_scaling_shape(args.size, world)
    nvox = shape[0] * shape[1] * shape[2]
    stepper = DistributedCahnHilliardIMEX(shape, (1.0, 1.0, 1.0), CH["dt"], CH["eps"], CH["D"],
                                          CH["A"], device=dev, transport=args.transport, copier=args.copier,
                                          scatter_ctas=args.scatter_ctas, mid_chunks=args.mid_chunks,
                                          overlap_chunks=args.overlap_chunks, p2p_ctas=args.p2p_ctas)
    gen = torch.Generator(device=dev).manual_seed(rank)
    u0 = 0.5 + 0.1 * torch.rand(stepper.slab.local_shape, device=dev, generator=gen)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(ms):
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    u = u0
    for _ in range(max(args.warmup, 3)):
        u = stepper.step(u)
    m_start = stepper.total_mass(u0)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        u = stepper.step(u)
    e1.record()
    barrier()
    launches = _native.launch_count() - l0
    ms = reduce_max(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else {}
    mass_drift = abs(stepper.total_mass(u) - m_start) / abs(m_start)

    e2e_steps = max(6, min(args.steps, 20))
    h_in = torch.empty(stepper.slab.local_shape, dtype=torch.float32, device="cpu", pin_memory=True)
    h_in.copy_(u0)
    h_outs = [torch.empty(stepper.slab.local_shape, dtype=torch.float32, device="cpu", pin_memory=True)
              for _ in range(2)]
    e2e_pipelined(stepper.step, h_in, h_outs, dev, 3)
    barrier()
    ms_e2e = reduce_max(e2e_pipelined(stepper.step, h_in, h_outs, dev, e2e_steps))
    barrier()
    if rank == 0:
        peak, peak_src = measured_peaks()
        ms_step = ms / args.steps
        gbs = B_ALG_STEP * nvox / world / (ms_step * 1e-3) / 1e9     # per GPU
        slab_bytes = nvox // world * 4
        line = {
            "metric": METRIC, "value": nvox * args.steps / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"CH IMEX {shape[0]}x{shape[1]}x{shape[2]} fp32 periodic dt=0.1 "
                                   f"({args.size}^3 voxels per GPU)", **CH, "fft_backend": "native",
                       "parallelism": f"x-slab over {world} GPUs: 2-plane halos ("
                                      + ("DMA writes into the neighbours' symmetric-memory slots" if args.transport == "ce"
                                         else "NCCL send/recv") + ") + slab<->pencil "
                                      + {"p2p": "transposes fused into the FFT passes as NVLink peer stores (symmetric memory)",
                                         "ce": "transposes as DMA-engine copies between symmetric-memory block buffers, "
                                               "pipelined against the kernels of the next chunk",
                                         "nccl": "NCCL all-to-all"}[args.transport],
                       "l2": "slab (%.0f MB) larger than L2 (126 MB), no flush needed" % (slab_bytes / 1e6)},
            "clocks": clocks,
            "e2e": {"value": nvox * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": slab_bytes * world, "d2h_bytes_per_step": slab_bytes * world,
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "note": "per rank: pinned host slab -> H2D -> distributed step -> D2H, every step; "
                            "legs of consecutive steps overlap on three streams"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "whole step per GPU (60 B/voxel)", "achieved": gbs,
                         "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                         "peak_source": peak_src},
            "cpu_baseline": None, "mass_drift": mass_drift,
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fft", default="auto", choices=["auto", "cufft", "native"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--p2p-ctas", type=int, default=148,
                    help="grid cap of the NVLink-bound peer-store launches (0 = fill the GPU)")
    ap.add_argument("--overlap-chunks", type=int, default=4,
                    help="multi-GPU p2p: x chunks pipelined on two streams in the forward half")
    ap.add_argument("--copier", default=None, choices=["kernel", "dma"],
                    help="multi-GPU 'ce' transport: blocks moved by one scatter kernel per chunk or by DMA copies")
    ap.add_argument("--mid-chunks", type=int, default=0,
                    help="multi-GPU 'ce' transport: chunks of the x pass (0 = same as --overlap-chunks)")
    ap.add_argument("--scatter-ctas", type=int, default=0, help="CTAs per peer of the scatter kernel (default 8)")
    ap.add_argument("--transport", default="ce", choices=["p2p", "nccl", "ce"],
                    help="multi-GPU transposes: fused peer stores over NVLink, or NCCL all-to-all")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
