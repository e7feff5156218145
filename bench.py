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


def workload_string(size, world):
    """One string for both arms (the driver compares them)."""
    if world == 1:
        return f"CH IMEX {size}^3 fp32 periodic dt=0.1"
    sh = weak_scaling_shape(size, world)
    return f"CH IMEX {sh[0]}x{sh[1]}x{sh[2]} fp32 periodic dt=0.1 ({size}^3 voxels per GPU)"


def ncu_traffic(kernel_name, n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel from the
    committed `ncu --set full` capture (profiles/r02_ncu_traffic.json, written by
    scripts/ncu_traffic.py from the raw export).  None if the capture has no such kernel."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
            tab = json.load(f)
        if int(tab.get("size", 0)) != n:
            return None, None
        for key, v in tab["kernels"].items():
            if key in kernel_name or kernel_name.split(" ")[0] == v.get("bench_name"):
                return float(v["dram_bytes"]), tab.get("source")
    except Exception:
        pass
    return None, None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML from a thread of this process, every
    few milliseconds, so that even a 30 ms timed region is covered by samples (an `nvidia-smi
    -lms` child needs ~0.5 s to start and cannot see it).  `mark()` brackets the timed region;
    samples taken while the GPU is busy before / after it (warm-up steps, clock-probe steps: the
    latter show the sustained clock under the power cap) are reported next to the in-region ones."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}

    def __init__(self, gpu_index=0, period_s=0.002):
        self.gpu, self.period = gpu_index, period_s
        self.samples = []          # (t, sm_mhz, reasons_mask, power_w)
        self.marks = []
        self._stop = threading.Event()
        self._thread = None
        self.error = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES if it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            try:
                ids = [int(x) for x in vis.split(",") if x.strip() != ""]
                if ids:
                    idx = ids[self.gpu]
            except (ValueError, IndexError):
                pass
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop.is_set():
                try:
                    mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    mask = int(get_reasons(h))
                    try:
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                    except Exception:
                        pw = None
                    self.samples.append((time.perf_counter(), mhz, mask, pw))
                except Exception as exc:      # keep sampling
                    self.error = repr(exc)
                time.sleep(self.period)
        except Exception as exc:
            self.error = repr(exc)

    def start(self):
        self.max_mhz = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def mark(self):
        self.marks.append(time.perf_counter())

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": getattr(self, "max_mhz", None), "reasons": [],
               "samples": 0, "samples_in_timed_region": 0, "sm_mhz_timed_region": None,
               "source": "NVML, in-process thread, %.0f ms period; sampled from the warm-up steps "
                         "through the timed region to the clock-probe steps after it" % (self.period * 1e3)}
        if self.error:
            out["error"] = self.error
        if not self.samples:
            return out
        t0, t1 = (self.marks + [None, None])[:2]
        inside = [x for x in self.samples if t0 is not None and t1 is not None and t0 <= x[0] <= t1]
        mask = 0
        for x in self.samples:
            mask |= x[2]
        pw = [x[3] for x in self.samples if x[3] is not None]
        out.update(sm_mhz=statistics.median(x[1] for x in self.samples), samples=len(self.samples),
                   sm_mhz_min=min(x[1] for x in self.samples),
                   reasons=sorted(n for b, n in self.REASONS.items() if mask & b),
                   samples_in_timed_region=len(inside), power_w_max=max(pw) if pw else None)
        if inside:
            out["sm_mhz_timed_region"] = statistics.median(x[1] for x in inside)
            m = 0
            for x in inside:
                m |= x[2]
            out["reasons_timed_region"] = sorted(n for b, n in self.REASONS.items() if m & b)
        return out


def busy_for(step_fn, u, seconds, dev):
    """Keep the GPU busy with untimed steps for about `seconds` (clock ramp-up / clock probe)."""
    import torch
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        for _ in range(20):
            u = step_fn(u)
        torch.cuda.synchronize(dev)
    return u


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

def time_reference(problem, size, steps, warmup, device="cpu", jit=False, budget_s=0.0, timeout_s=600):
    """oracle/time_reference.py in a subprocess: the UNMODIFIED reference (from baseline/_ref or
    /root/reference, kind "reference") or, where neither exists, the oracle port (kind "port")."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "time_reference.py"), "--problem", problem,
           "--size", str(size), "--steps", str(steps), "--warmup", str(warmup), "--device", device]
    if jit:
        cmd.append("--jit")
    if budget_s:
        cmd += ["--budget-s", str(budget_s)]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        lines = [l for l in res.stdout.strip().splitlines() if l.startswith("{")]
        if not lines:
            return {"kind": "failed", "error": (res.stderr or res.stdout)[-300:]}
        return json.loads(lines[-1])
    except subprocess.TimeoutExpired:
        return {"kind": "timeout", "error": f"no result within {timeout_s} s"}
    except Exception as exc:
        return {"kind": "failed", "error": repr(exc)[:300]}


def cpu_baseline_record(size, steps, warmup, budget_s=40.0):
    r = time_reference("ch", size, steps, warmup, "cpu", budget_s=budget_s)
    if "vox_per_s" not in r:
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": r.get("kind", "failed"),
                "sample": "failed: " + str(r.get("error"))}
    what = ("unmodified reference (evoxels PseudoSpectralIMEX.step, torch CPU eager, through oracle/ref_shim.py)"
            if r["kind"] == "reference" else "oracle port of the reference (torch CPU eager)")
    return {"value": r["vox_per_s"], "unit": UNIT, "cores": r["threads"], "kind": r["kind"],
            "sample": f"{r['steps']} steps of {size}^3 after {warmup} warm-up, {what}",
            "s_per_step": r["s_per_step"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    # bounded sample: the largest cube whose (K+W) steps fit ~150 s on these cores
    probe = time_reference("ch", 128, 1, 1, "cpu")
    t128 = probe.get("s_per_step", 1.0)
    total = args.steps + args.warmup
    size = 128
    for cand, factor in ((512, 64 * 3.0), (256, 8 * 2.5)):   # voxel ratio x cache penalty
        if t128 * factor * total <= 150.0:
            size = cand
            break
    r = time_reference("ch", size, args.steps, args.warmup, "cpu")
    vps, spstep = r.get("vox_per_s"), r.get("s_per_step", 0.0)
    kind = r.get("kind", "failed")
    line = {
        "impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": spstep * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_string(args.size, world), **CH,
                   "sample": f"{size}^3 cube of that workload on the host cores (the CPU path cannot "
                             f"hold K+W steps of the full grid within minutes)" if size != args.size or world > 1
                             else "full workload"},
        "cpu_baseline": {"value": vps, "unit": UNIT, "cores": r.get("threads"), "kind": kind,
                         "sample": f"{args.steps} steps of {size}^3 after {args.warmup} warm-up, "
                                   + ("unmodified reference, torch CPU eager" if kind == "reference"
                                      else "oracle port, torch CPU eager")},
        "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if "error" in r:
        line["error"] = r["error"]
    print(json.dumps(line), flush=True)


# ---- other BASELINE configs, timed outside the headline region -------------------------------
def extra_configs(dev):
    """config 1 (README 100^3), config 3 (Allen-Cahn 1024^3 Neumann, one GPU) and config 5
    (inversion: forward + backward through 100 steps at 256^3) - one number each."""
    import torch
    import evoxels_b200 as evo
    from evoxels_b200.problem_definition import CahnHilliard, TwoPhaseAllenCahn
    from evoxels_b200.timesteppers import ForwardEuler, PseudoSpectralIMEX
    from evoxels_b200.voxelgrid import VoxelGridTorch
    peak, _ = measured_peaks()
    out = {}

    def events(fn):
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b)

    try:   # config 1
        n = 100
        vf = evo.VoxelFields((n, n, n), (float(n),) * 3)
        vg = VoxelGridTorch(vf.grid_info(), device=str(dev))
        ts = PseudoSpectralIMEX(CahnHilliard(vg, eps=3.0, D=1.0), 0.1)
        u = 0.5 + 0.1 * torch.rand((1, n, n, n), device=dev)
        state = {"u": u}

        def run(k):
            v = state["u"]
            for _ in range(k):
                v = ts.step(0.0, v)
            state["u"] = v
        run(20)
        ms = events(lambda: run(1000))
        rec = {"us_per_step": ms, "steps": 1000, "voxel_updates_per_s": n ** 3 * 1000 / (ms * 1e-3),
               "fft_backend": next(iter(ts._plans.values())).backend_name,
               "note": "eager step loop (no CUDA graph), 1000 steps as in the README example"}
        del ts, u, state
        try:      # the README call itself: run_cahn_hilliard_solver (CUDA-graph replay, 10 frames)
            import numpy as np
            vf2 = evo.VoxelFields((n, n, n), (float(n),) * 3)
            vf2.add_field("c", (0.5 + 0.1 * np.random.default_rng(0).random((n, n, n))).astype(np.float32))
            evo.run_cahn_hilliard_solver(vf2, "c", backend="torch", device=str(dev), time_increment=0.1,
                                         frames=10, max_iters=1000, verbose=False)          # warm-up
            vf2.add_field("c", (0.5 + 0.1 * np.random.default_rng(0).random((n, n, n))).astype(np.float32))
            sol = evo.run_cahn_hilliard_solver(vf2, "c", backend="torch", device=str(dev), time_increment=0.1,
                                               frames=10, max_iters=1000, verbose=False)
            rec["us_per_step_readme_call"] = sol.computation_time / 1000 * 1e6
            rec["readme_call"] = "run_cahn_hilliard_solver(..., frames=10, max_iters=1000), wall clock incl. frame export"
        except Exception as exc:
            rec["readme_call_error"] = repr(exc)[:200]
        out["config1_ch_100^3_readme"] = rec
    except Exception as exc:
        out["config1_ch_100^3_readme"] = {"error": repr(exc)[:200]}
    try:   # config 3
        n = 1024
        vf = evo.VoxelFields((n, n, n), (float(n),) * 3)
        vg = VoxelGridTorch(vf.grid_info(), device=str(dev))
        ts = ForwardEuler(TwoPhaseAllenCahn(vg), 0.05)
        state = {"u": torch.rand((1, n, n, n), device=dev)}

        def run(k):
            v = state["u"]
            for _ in range(k):
                v = ts.step(0.0, v)
            state["u"] = v
        run(3)
        ms = events(lambda: run(10)) / 10
        out["config3_ac_1024^3_neumann_1gpu"] = {
            "ms_per_step": ms, "voxel_updates_per_s": n ** 3 / (ms * 1e-3),
            "achieved_GBs_at_8B_per_voxel": 8 * n ** 3 / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": 8 * n ** 3 / (ms * 1e-3) / 1e9 / peak,
            "finite": bool(torch.isfinite(state["u"][0, ::64]).all())}
        del ts, state
        torch.cuda.empty_cache()
    except Exception as exc:
        out["config3_ac_1024^3_neumann_1gpu"] = {"error": repr(exc)[:200]}
    try:   # config 5
        n, steps = 256, 100
        vf = evo.VoxelFields((n, n, n), (float(n),) * 3)
        vg = VoxelGridTorch(vf.grid_info(), device=str(dev))
        u0 = 0.5 + 0.1 * torch.rand((1, n, n, n), device=dev)
        obs_at = {steps // 3, 2 * steps // 3, steps}
        with torch.no_grad():
            ts = PseudoSpectralIMEX(CahnHilliard(vg, eps=3.0, D=1.0), 0.1)
            v, obs = u0, {}
            for i in range(1, steps + 1):
                v = ts.step(0.0, v)
                if i in obs_at:
                    obs[i] = v.clone()

        def fwd_bwd():
            D = torch.tensor(2.0, dtype=torch.float64, device=dev, requires_grad=True)
            eps = torch.tensor(2.0, dtype=torch.float64, device=dev, requires_grad=True)
            tsg = PseudoSpectralIMEX(CahnHilliard(vg, eps=eps, D=D), 0.1)
            w, loss = u0.clone().requires_grad_(True), 0.0
            for i in range(1, steps + 1):
                w = tsg.step(0.0, w)
                if i in obs_at:
                    loss = loss + ((w - obs[i]) ** 2).sum()
            return torch.autograd.grad(loss, (D, eps))
        fwd_bwd()
        ms = events(fwd_bwd)
        out["config5_inversion_256^3_x100"] = {
            "ms_forward_plus_backward": ms, "ms_per_step": ms / steps,
            "achieved_GBs_at_160B_per_voxel": 160 * n ** 3 * steps / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": 160 * n ** 3 * steps / (ms * 1e-3) / 1e9 / peak}
    except Exception as exc:
        out["config5_inversion_256^3_x100"] = {"error": repr(exc)[:200]}
    torch.cuda.empty_cache()
    return out


def reference_gpu_record(n):
    """The reference's own GPU path on this B200 (eager, and torch.compile(step) as
    evoxels/solvers.py:64-70 does with jit=True): the number the kernels have to beat."""
    out = {}
    for tag, jit, tmo in (("eager", False, 240), ("torch_compile", True, 420)):
        r = time_reference("ch", n, 5, 2, "cuda", jit=jit, timeout_s=tmo)
        if "s_per_step" in r:
            out[tag] = {"ms_per_step": r["s_per_step"] * 1e3, "voxel_updates_per_s": r["vox_per_s"],
                        "kind": r["kind"], "peak_mem_GB": r.get("peak_mem_GB")}
        else:
            out[tag] = {"kind": r.get("kind"), "error": r.get("error")}
    out["cite"] = "evoxels/solvers.py:64-70, voxelgrid.py:203-207 (step = timestepper.step, optionally torch.compile'd)"
    return out


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
    step1 = lambda v: ts.step(0.0, v)      # noqa: E731
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    u = u0
    for _ in range(max(args.warmup, 3)):
        u = ts.step(0.0, u)
    plan = next(iter(ts._plans.values()))
    barrier()
    l0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark()
    e0.record()
    for _ in range(args.steps):
        u = ts.step(0.0, u)
    e1.record()
    barrier()
    sampler.mark()
    launches = _native.launch_count() - l0
    ms = e0.elapsed_time(e1)
    u_end = u
    busy_for(step1, u, 0.7, dev)           # clock probe: same kernels, >= 1 s of samples in total
    clocks = sampler.stop() if rank == 0 else {}
    u = u_end
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
        def one_pass(w):
            return lambda: plan.native_pass(w, u0[0], rhs_buf[0], out_buf[0], vg.spacing, CH["dt"], coef, 2)
        try:
            one_pass(5)()
            chained = True
        except Exception:
            chained = False
        if chained:      # one persistent kernel per z/y pair, the pair's intermediate stays in L2
            passes = [(5, "fft_zy_forward (chained ZPass + StridedLine FWD)", 4 + 4 * half),
                      (2, "fft_x_fwd*filter*inv (StridedLine XMID)", 8 * half),
                      (6, "fft_yz_inverse+u (chained StridedLine INV + ZPass)", 4 * half + 8)]
        else:
            passes = [(0, "fft_z_forward (ZPass fwd)", 4 + 4 * half), (1, "fft_y_forward (strided FWD)", 8 * half),
                      (2, "fft_x_fwd*filter*inv (strided XMID)", 8 * half),
                      (3, "fft_y_inverse (strided INV)", 8 * half), (4, "fft_z_inverse+u (ZPass inv)", 4 * half + 8)]
        for w, nm, bb in passes:
            kernels[nm] = (timed(one_pass(w)), bb)
    else:
        kernels["spectral_apply (cuFFT + filter + add kernels)"] = (
            timed(lambda: plan.apply(u0[0], rhs_buf[0], out_buf[0], vg.spacing, CH["dt"], coef, 2)), 52.0)
    dom = max(kernels, key=lambda k: kernels[k][0])
    dom_ms, dom_bpv = kernels[dom]
    dom_gbs = dom_bpv * nvox / (dom_ms * 1e-3) / 1e9
    gbs_step = B_ALG_STEP * nvox / (ms_step * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak, "unit": "GB/s",
        "frac": dom_gbs / peak, "traffic": ncu_traffic(dom, n)[0],
        "traffic_source": ncu_traffic(dom, n)[1],
        "peak_source": peak_src, "ms": dom_ms,
        "bytes_per_voxel": dom_bpv,
        "step": {"bytes_per_voxel": B_ALG_STEP, "achieved": gbs_step, "frac": gbs_step / peak,
                 "frac_of_8TBs": gbs_step / 8000.0, "ms": ms_step},
        "kernels": {k: {"ms": v[0], "bytes_per_voxel": v[1],
                        "achieved_GBs": v[1] * nvox / (v[0] * 1e-3) / 1e9,
                        "frac": v[1] * nvox / (v[0] * 1e-3) / 1e9 / peak} for k, v in kernels.items()},
        # second roofline of these kernels (not measured live: one ncu capture per round)
        "shared_memory_pipe": {
            "note": "the FFT kernels are bound by shared-memory wavefronts, not by HBM: every radix-8 "
                    "exchange sends each spectrum point through shared memory twice",
            "pct_of_peak": {"fft_chain forward": 77.0, "x pass (512-point lines)": 76.9,
                            "fft_chain inverse": 58.1, "ch_rhs": 41.4},
            "metric": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "source": "profiles/r02_ncu_smem_pipe.txt (ncu --set full, 512^3, one launch each)"},
    }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- beside it, outside the timed regions (rank 0, N = 1 only; bounded samples) --------------
    extras = None if args.no_extras else extra_configs(dev)
    del rhs_buf, out_buf, h_in, h_outs
    torch.cuda.empty_cache()
    cpu = None
    ref_gpu = None
    if not args.no_cpu:
        cs = 512 if n >= 512 else n
        cpu = cpu_baseline_record(cs, 3, 1)
        ref_gpu = reference_gpu_record(n)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(n, 1), **CH,
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
        "reference_gpu": ref_gpu,
        "other_configs": extras,
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


def distributed_parity(world, rank, dev, transport):
    """Distributed step == single-GPU step on the same 256^3 global field (every rank computes
    the single-GPU step of the whole field itself and compares its own slab), for every
    transport, plus bit equality across the transports.  Runs before the timed region; the
    caller exits non-zero when rel-L2 exceeds 1e-6."""
    import torch
    import torch.distributed as dist
    from evoxels_b200 import _native
    from evoxels_b200.distributed import DistributedCahnHilliardIMEX
    shape = (256, 256, 256)
    gen = torch.Generator(device=dev).manual_seed(1234)      # same stream on every rank
    ug = 0.5 + 0.1 * torch.rand(shape, device=dev, generator=gen)
    plan = _native.ImexPlan(shape, torch.float32, dev, _native.FFT_NATIVE)
    ref = torch.empty_like(ug)
    plan.ch_step(ug, ref, (1.0, 1.0, 1.0), CH["dt"], CH["eps"], CH["D"], CH["A"])
    outs, rec = {}, {}
    for tr in dict.fromkeys([transport, "ce", "p2p", "nccl"]):
        try:
            st = DistributedCahnHilliardIMEX(shape, (1.0, 1.0, 1.0), CH["dt"], CH["eps"], CH["D"], CH["A"],
                                             device=dev, transport=tr)
            ul = st.slab.take(ug).contiguous()
            out = st.step(ul)
            torch.cuda.synchronize(dev)
            rl = st.slab.take(ref)
            sums = torch.stack([((out - rl).double() ** 2).sum(), (rl.double() ** 2).sum(),
                                ((rl - ul).double() ** 2).sum()])
            dist.all_reduce(sums)
            rec[tr] = {"rel_l2": float((sums[0] / sums[1]).sqrt()),
                       "update_rel_l2": float((sums[0] / sums[2]).sqrt())}
            outs[tr] = out
            del st
        except Exception as exc:
            rec[tr] = {"error": repr(exc)[:200]}
    keys = [k for k in outs]
    same = torch.tensor([1.0 if keys and all(torch.equal(outs[keys[0]], outs[k]) for k in keys[1:]) else 0.0],
                        device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    main = rec.get(transport, {})
    return {"grid": "256^3 global", "vs": "single-GPU evx_ch_imex_step_f32 of the whole field",
            "rel_l2": main.get("rel_l2"), "update_rel_l2": main.get("update_rel_l2"),
            "bit_equal_across_transports": bool(same.item() == 1.0) and len(keys) > 1,
            "transports": rec}


def distributed_ac_config3(world, rank, dev):
    """BASELINE config 3: Allen-Cahn forward Euler, 1024^3, Neumann, x slabs over the ranks."""
    import torch
    import torch.distributed as dist
    from evoxels_b200.distributed import DistributedAllenCahnEuler
    n = 1024
    st = DistributedAllenCahnEuler((n, n, n), (1.0, 1.0, 1.0), 0.05, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    phi = torch.rand(st.slab.local_shape, device=dev, generator=gen)
    for _ in range(3):
        phi = st.step(phi)
    dist.barrier()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        phi = st.step(phi)
    b.record()
    dist.barrier()
    torch.cuda.synchronize(dev)
    t = torch.tensor([a.elapsed_time(b) / 20], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    peak, _ = measured_peaks()
    return {"ms_per_step": ms, "voxel_updates_per_s": n ** 3 / (ms * 1e-3),
            "achieved_GBs_per_gpu_at_8B_per_voxel": 8 * n ** 3 / world / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak_per_gpu": 8 * n ** 3 / world / (ms * 1e-3) / 1e9 / peak,
            "finite": bool(torch.isfinite(phi[::16]).all())}


def config4_record(world, rank, dev):
    """BASELINE config 4 (opt-in, `--config4`): Cahn-Hilliard 2048^3 on all ranks of this run, then
    the same grid on rank 0 alone (137 GB of HBM) for the scaling ratio the north star states."""
    import torch
    import torch.distributed as dist
    from evoxels_b200 import _native
    from evoxels_b200.distributed import DistributedCahnHilliardIMEX
    n = 2048
    rec = {"grid": f"{n}^3", "n_gpus": world}
    st = DistributedCahnHilliardIMEX((n, n, n), (1.0, 1.0, 1.0), CH["dt"], CH["eps"], CH["D"], CH["A"],
                                     device=dev, transport="ce")
    gen = torch.Generator(device=dev).manual_seed(rank)
    u = torch.empty(st.slab.local_shape, device=dev)
    for i in range(0, u.shape[0], 32):
        u[i:i + 32] = 0.5 + 0.1 * torch.rand((min(32, u.shape[0] - i), n, n), device=dev, generator=gen)
    m0 = st.total_mass(u)
    for _ in range(2):
        u = st.step(u)
    dist.barrier()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        u = st.step(u)
    b.record()
    dist.barrier()
    torch.cuda.synchronize(dev)
    t = torch.tensor([a.elapsed_time(b) / 5], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    rec["ms_per_step"] = float(t.item())
    rec["voxel_updates_per_s"] = n ** 3 / (rec["ms_per_step"] * 1e-3)
    rec["mass_drift"] = abs(st.total_mass(u) - m0) / abs(m0)
    del st, u
    torch.cuda.empty_cache()
    dist.barrier()
    if rank == 0:
        try:
            plan = _native.ImexPlan((n, n, n), torch.float32, dev, _native.FFT_NATIVE)
            v = torch.empty((n, n, n), device=dev)
            for i in range(0, n, 32):
                v[i:i + 32] = 0.5 + 0.1 * torch.rand((32, n, n), device=dev, generator=gen)
            w = torch.empty_like(v)
            for _ in range(2):
                plan.ch_step(v, w, (1, 1, 1), CH["dt"], CH["eps"], CH["D"], CH["A"])
                v, w = w, v
            torch.cuda.synchronize(dev)
            a.record()
            for _ in range(3):
                plan.ch_step(v, w, (1, 1, 1), CH["dt"], CH["eps"], CH["D"], CH["A"])
                v, w = w, v
            b.record()
            torch.cuda.synchronize(dev)
            one = a.elapsed_time(b) / 3
            rec["one_gpu_ms_per_step"] = one
            rec["speedup_vs_one_gpu"] = one / rec["ms_per_step"]
            rec["one_gpu_peak_mem_GB"] = torch.cuda.max_memory_allocated(dev) / 1e9
            del plan, v, w
        except Exception as exc:
            rec["one_gpu_error"] = repr(exc)[:200]
        torch.cuda.empty_cache()
    dist.barrier()
    return rec


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

    parity = None
    if not args.no_extras:
        parity = distributed_parity(world, rank, dev, args.transport)
        if parity["rel_l2"] is None or not parity["rel_l2"] <= 1e-6:
            if rank == 0:
                print(json.dumps({"error": "distributed step does not match the single-GPU step", "parity": parity}),
                      flush=True)
            dist.destroy_process_group()
            raise SystemExit(3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    u = u0
    for _ in range(max(args.warmup, 3)):
        u = stepper.step(u)
    m_start = stepper.total_mass(u0)
    barrier()
    l0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark()
    e0.record()
    for _ in range(args.steps):
        u = stepper.step(u)
    e1.record()
    barrier()
    sampler.mark()
    launches = _native.launch_count() - l0
    ms = reduce_max(e0.elapsed_time(e1))
    u_end = u
    for _ in range(200):                    # clock probe
        u = stepper.step(u)
    torch.cuda.synchronize(dev)
    clocks = sampler.stop() if rank == 0 else {}
    u = u_end
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
    ac3 = None
    if not args.no_extras:
        del h_in, h_outs, u, u_end
        torch.cuda.empty_cache()
        try:
            ac3 = distributed_ac_config3(world, rank, dev)
        except Exception as exc:
            ac3 = {"error": repr(exc)[:200]}
    stepper_direct = int(getattr(stepper.ops, "direct_peers", 0))
    cfg4 = None
    if args.config4 or (world == 8 and not args.no_extras):
        del stepper
        torch.cuda.empty_cache()
        try:
            cfg4 = config4_record(world, rank, dev)
        except Exception as exc:
            cfg4 = {"error": repr(exc)[:200]}
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
            "config": {"workload": workload_string(args.size, world), **CH, "fft_backend": "native",
                       "parallelism": f"x-slab over {world} GPUs: 2-plane halos ("
                                      + ("DMA writes into the neighbours' symmetric-memory slots" if args.transport == "ce"
                                         else "NCCL send/recv") + ") + slab<->pencil "
                                      + {"p2p": "transposes fused into the FFT passes as NVLink peer stores (symmetric memory)",
                                         "ce": "transposes as DMA-engine copies between symmetric-memory block buffers, "
                                               "pipelined against the kernels of the next chunk"
                                               + (f"; the blocks of {stepper_direct} of the {world - 1} peers leave the FFT passes "
                                                  "as TMA stores into the peer's buffer over NVLink" if stepper_direct else ""),
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
            "cpu_baseline": None, "mass_drift": mass_drift, "parity": parity,
            "other_configs": {"config3_ac_1024^3_neumann": ac3, "config4_ch_2048^3": cfg4},
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline and reference-on-GPU legs")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 1 / 3 / 5 sub-records")
    ap.add_argument("--config4", action="store_true",
                    help="N > 1: also time Cahn-Hilliard 2048^3 on all ranks and on rank 0 alone (config 4); "
                         "on by default at N = 8, the configuration BASELINE.json names")
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
