"""Time the UNMODIFIED reference's step on this box (CPU cores or its own GPU path).

TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT: executed only by bench.py's `cpu_baseline`,
`reference_gpu` and `--impl reference` legs, always in a subprocess (the reference's
VoxelGridTorch calls torch.set_default_device globally, evoxels/voxelgrid.py:178).

    python oracle/time_reference.py --problem ch --size 512 --steps 2 --warmup 1 \
        --device cpu|cuda [--jit] [--budget-s 60]

The reference sources are looked up in /root/reference (build container) and then in
baseline/_ref (a `pip install --target` of the unmodified reference, which travels to the GPU
box); when neither exists the line says kind = "port" and the restatement in evx_oracle.py is
timed instead (CPU only).  The stepping sequence is the reference's own
(evoxels/solvers.py:52-72,192-208): problem_cls(vg, **kw), timestepper_cls(problem, dt).step,
optionally torch.compile(step), then u = step(t, u) in a loop.
Prints one JSON line: {"kind", "device", "jit", "size", "steps", "s_per_step", "vox_per_s",
"threads", "note"}.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def find_reference():
    for cand in (os.environ.get("EVOXELS_REFERENCE"), "/root/reference",
                 os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "evoxels")):
            return cand
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problem", default="ch", choices=["ch", "ac"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--jit", action="store_true")
    ap.add_argument("--budget-s", type=float, default=0.0,
                    help="stop stepping once this much time has been spent (>= 1 timed step)")
    args = ap.parse_args()

    import numpy as np
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    n = args.size
    shape = (n, n, n)
    cuda = args.device.startswith("cuda")
    out = {"device": args.device, "jit": bool(args.jit), "size": n, "problem": args.problem,
           "threads": torch.get_num_threads()}
    root = find_reference()
    if root is None:
        # no reference sources on this box: time the restatement (same torch calls in the same
        # order, oracle/evx_oracle.py) - on the GPU too, where it stands for the reference's
        # own eager-CUDA path (evoxels/solvers.py:64-70, voxelgrid.py:203-207)
        from oracle import evx_oracle as O
        out["kind"] = "port"
        if args.problem == "ch":
            orc = O.CHOracle(shape, (1.0, 1.0, 1.0), 0.1, 3.0, 1.0, 0.25)
            orc.prefac = orc.prefac.to(args.device)
            u = O.noise_field(shape, seed=0).to(args.device)
        else:
            orc = O.ACOracle(shape, (1.0, 1.0, 1.0), 0.05)
            u = O.noise_field(shape, seed=1, lo=0.0, amp=1.0).to(args.device)
        step = lambda t, v: orc.step(v)      # noqa: E731
        if args.jit:
            step = torch.compile(step)
    else:
        os.environ["EVOXELS_REFERENCE"] = root
        from oracle.ref_shim import load_reference
        ref = load_reference()
        out["kind"] = "reference"
        out["reference_root"] = root
        vf = ref.voxelfields.VoxelFields(shape, tuple(float(s) for s in shape))
        vg = ref.voxelgrid.VoxelGridTorch(vf.grid_info(), "float32", args.device)
        if args.problem == "ch":
            prob = ref.problem_definition.CahnHilliard(vg, eps=3.0, D=1.0, A=0.25)
            ts = ref.timesteppers.PseudoSpectralIMEX(prob, 0.1)
            a = 0.5 + 0.1 * np.random.default_rng(0).random(shape).astype(np.float32)
        else:
            prob = ref.problem_definition.TwoPhaseAllenCahn(vg)
            ts = ref.timesteppers.ForwardEuler(prob, 0.05)
            a = np.random.default_rng(1).random(shape).astype(np.float32)
        u = vg.init_scalar_field(a)
        step = ts.step
        if args.jit:
            step = torch.compile(step)       # evoxels/solvers.py:64-70

    def sync():
        if cuda:
            torch.cuda.synchronize()

    t_begin = time.perf_counter()
    try:
        for i in range(args.warmup):
            u = step(0.0, u)
        sync()
        t0 = time.perf_counter()
        done = 0
        for i in range(args.steps):
            u = step(0.0, u)
            done += 1
            if args.budget_s and not cuda and time.perf_counter() - t_begin > args.budget_s:
                break
        sync()
        dt = (time.perf_counter() - t0) / max(done, 1)
        out.update(steps=done, warmup=args.warmup, s_per_step=dt, vox_per_s=n ** 3 / dt,
                   finite=bool(torch.isfinite(u).all()))
        if cuda:
            out["peak_mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
    except Exception as exc:   # e.g. torch.compile not usable on this box
        out.update(kind=out.get("kind", "reference"), error=repr(exc)[:300])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
