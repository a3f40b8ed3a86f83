#!/usr/bin/env python
"""bench.py -- G body-body interactions/s of the brute-force step (force + fused integrator).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n N] [--precision 32|64]

One "step" = one pass of the hot path over all N bodies: all N(N-1) ordered interactions plus
the per-body update (v += F/m dt; x += v dt).  The metric counts ORDERED interactions delivered
(SURVEY 8d): the pair-symmetric pass evaluates each unordered pair once and feeds both bodies,
exactly like the reference's own j > i loop (methods.cpp:18-39), and is credited with both.  Workload at every GPU count = BASELINE.json
configs[4]: 3D, N = 1,048,576 bodies in a uniform cube (strong scaling: N is fixed, targets are
sharded over the GPUs).  For --gpus N > 1 launch under torchrun, one rank per GPU.

Product arm: libnb200.so through its C ABI (include/nb200.h).  `value` is timed with CUDA
events recorded by the library on its compute stream around each step, inputs resident in HBM,
MAX over ranks; `e2e` goes through the public host-buffer API (upload from pinned host memory +
step + download, every step) by wall clock.  Reference arm (--impl reference): the reference's
own brute-force code compiled unmodified (oracle/_ref) -- or the oracle port when that library
is absent -- on this box's host cores, on a bounded sample of the same workload.

Nothing here imports oracle/ on the product path: only `cpu_baseline` and `--impl reference` do.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

# rank 0 prints exactly ONE JSON line on stdout.  NCCL_DEBUG is left entirely to the caller: unset, NCCL prints nothing
# (not even its version banner); a caller that asks for INFO (the driver's communicator check) gets it untouched.

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

N_DEFAULT = 1 << 20
DIM = 3
DT = 1e-3
SEED = 42 + 5
# BASELINE.json configs (SURVEY 8d inputs: seeded, uniform cube / Plummer).  dt: the reference defines none
# (SURVEY F4); c2/c3 use a step that resolves the closest pairs of the uniform sets so that the 100-step
# energy drift is a statement about the integrator, not about two bodies flung apart in one step.
CONFIGS = {
    "c2": {"dim": 3, "n": 16384, "dist": "cube", "steps": 100, "dt": 1e-4, "seed": 44},
    # c3: Poisson-uniform 2D points hold pairs that no fixed dt resolves (the closest ones sit at the cut-off radius
    # and are flung apart in a single step); the force parity uses the uniform SURVEY inputs, the 100 timed steps
    # and their energy drift a stratified ("jittered") uniform set of the same density
    "c3": {"dim": 2, "n": 65536, "dist": "cube", "steps": 100, "dt": 1e-5, "seed": 45, "steps_dist": "jittered"},
    "c4": {"dim": 3, "n": 262144, "dist": "plummer", "steps": 100, "dt": 1e-3, "seed": 46},
    "c5": {"dim": 3, "n": N_DEFAULT, "dist": "cube", "steps": 10, "dt": DT, "seed": SEED},
}
PARITY_TARGETS = 1024
FLOPS_PER_INTERACTION = 20          # north_star / GPU-Gems convention
SM_LANES_FP32 = 128                 # FP32 FMA lanes per SM (sm_100)
CPU_SAMPLE_N = 65536                # bounded CPU sample of the same distribution (~10-20 s of CPU work in total)


def interactions(n: int) -> float:
    return float(n) * float(n - 1)


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.idx)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self) -> dict:
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------- CPU legs
def cpu_reference_leg(n_sample: int, dim: int = DIM, cfg: dict | None = None, repeats: int = 1) -> dict:
    """Time the reference's own brute-force variants (oracle/_ref; else the oracle port) on this
    box's host cores, one force evaluation each (the reference's convention, utils.h:87-104)."""
    oracle = entry.load_oracle()
    pkg = entry.load_package()
    cfg = cfg or CONFIGS["c5"]
    bodies = make_bodies(pkg.generators, cfg)[:n_sample].copy()
    inter = interactions(n_sample)
    cores = os.cpu_count() or 1
    out = {"unit": "G interactions/s", "cores": cores,
           "sample": f"first {n_sample} bodies of the {dim}D {cfg['dist']} workload, one force evaluation per variant"}
    if oracle.have_ref():
        out["kind"] = "reference"
        thr = oracle.ref_threads()
        out["threads"] = {"omp": thr["omp"], "parlay": thr["parlay"]}
        variants = {}
        for v in ("omp_2", "omp_1", "parlay_1", "parlay_2"):
            best = min(oracle.ref_forces(bodies, v, want_forces=False)[1] for _ in range(repeats))
            variants[v] = round(inter / best / 1e9, 4)
        out["variants"] = variants
        out["value"] = max(variants.values())
        out["best_variant"] = max(variants, key=variants.get)
        out["cores"] = max(thr["omp"], thr["parlay"])
    else:
        out["kind"] = "port"
        t0 = time.perf_counter()
        oracle.forces(bodies)
        out["value"] = round(inter / (time.perf_counter() - t0) / 1e9, 4)
        out["cores"] = oracle.num_threads()
    return out


def run_reference_arm(args) -> None:
    """--impl reference: the reference's CPU implementation of the step on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the reference arm is the CPU implementation "with
    # all the host threads it can use", and only rank 0 runs: give it the whole box back (libgomp reads
    # the variable when the checker library is first loaded, below)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    oracle = entry.load_oracle()
    pkg = entry.load_package()
    cfg = dict(CONFIGS[args.config])
    if args.n:
        cfg["n"] = args.n
    args.n = cfg["n"]
    DIM, DT = cfg["dim"], cfg["dt"]
    n_s = min(args.n, CPU_SAMPLE_N)
    bodies = make_bodies(pkg.generators, cfg)[:n_s].copy()
    use_ref = oracle.have_ref()
    # the reference arm gets the reference's FASTEST brute-force variant on this host
    variant = "omp_2"
    if use_ref:
        speeds = {v: oracle.ref_forces(bodies, v, want_forces=False)[1]
                  for v in ("omp_1", "omp_2", "parlay_1", "parlay_2")}
        variant = min(speeds, key=speeds.get)

    def one_step(b):
        return oracle.ref_simulate(b, DT, 1, variant) if use_ref else oracle.simulate(b, DT, 1)

    for _ in range(args.warmup):
        bodies = one_step(bodies)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bodies = one_step(bodies)
    dt = time.perf_counter() - t0
    val = interactions(n_s) * args.steps / dt / 1e9
    cores = (oracle.ref_threads()["omp"] if use_ref else oracle.num_threads())
    line = {
        "impl": "reference", "metric": f"G body-body interactions/s (brute-force step, {DIM}D)", "value": round(val, 4),
        "unit": "G interactions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config}: brute-force {DIM}D {cfg['dist']}, N={args.n} (reference arm: CPU sample of {n_s} bodies)",
                   "n": args.n, "dim": DIM, "dt": DT},
        "cpu_baseline": {"value": round(val, 4), "unit": "G interactions/s", "cores": cores,
                         "kind": "reference" if use_ref else "port",
                         "sample": f"{args.steps} steps of brute_force_{variant}_n_body + update_body_* on the first "
                                   f"{n_s} bodies of the workload (methods.cpp:45-224,426-450; fastest of the "
                                   "reference's four parallel variants on this host), all host threads"},
        "e2e": {"value": round(val, 4), "unit": "G interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- ncu evidence
def ncu_traffic(n: int, prec: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the force kernel, per launch, from the committed
    `ncu --set full` capture of this workload (the newest profiles/**/*_summary.json, written by tools/ncu_summary.py)."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "**", f"*force_f{prec}_n{n}_*summary.json"), recursive=True)):
        try:
            k = json.load(open(f))["kernels"][0]
            best = (int(k["dram_traffic_bytes"]), os.path.relpath(f, ROOT))
        except Exception:
            continue
    return best if best else (None, None)


# ------------------------------------------------------------------------------------- host buffers
def shared_pinned_bodies(bodies, rank, world, torch, barrier):
    """The e2e leg's host array: page-locked; for world > 1 one /dev/shm mapping shared by all ranks."""
    rt = torch.cuda.cudart()
    if world == 1:
        host = torch.from_numpy(bodies.copy()).pin_memory()
        return host.numpy(), (lambda: None)
    path = f"/dev/shm/nb200_bench_{os.environ.get('MASTER_PORT', '0')}.f64"
    if rank == 0:
        mm = np.memmap(path, dtype=np.float64, mode="w+", shape=bodies.shape)
        mm[:] = bodies
        mm.flush()
    barrier()
    if rank != 0:
        mm = np.memmap(path, dtype=np.float64, mode="r+", shape=bodies.shape)
    rc = rt.cudaHostRegister(mm.ctypes.data, mm.nbytes, 0)
    if int(rc) != 0:
        raise RuntimeError(f"cudaHostRegister failed: {rc}")

    def release():
        rt.cudaHostUnregister(mm.ctypes.data)
        barrier()
        if rank == 0:
            os.unlink(path)

    return mm, release


# ------------------------------------------------------------------------------------- parity gate
def make_bodies(gen, cfg, dist=None):
    dist = dist or cfg["dist"]
    if dist == "plummer":
        return gen.plummer(cfg["n"], seed=cfg["seed"])
    if dist == "jittered":
        return gen.jittered_cube(cfg["n"], cfg["dim"], seed=cfg["seed"])
    return gen.uniform_cube(cfg["n"], cfg["dim"], seed=cfg["seed"])


def band_ok(err, kappa, prec):
    """The parity criterion of tests/test_gpu_parity.py on a set of (sampled) bodies."""
    if prec == 64:
        return bool(err.max() <= 1e-12)
    bound = entry.load_package().fp32_error_bound(kappa)
    return bool(np.all(err <= bound) and np.percentile(err, 99) <= 1e-5)


def parity_gate(pkg, oracle, D, ctx, bodies, dim, prec, dt, rank, world, local, bcast_ok):
    """After the timed region: a fresh force evaluation and ONE fused step of the context that was timed
    (same library, same launch path, all ranks), checked on PARITY_TARGETS sampled targets against the
    CPU oracle (checker only) and, for N > 1 GPUs, against a 1-GPU context of the same library on rank 0.
    The step is checked through the forces it implies: F = m (v1 - v0) / dt, and x1 = x0 + v1 dt."""
    gen = pkg.generators
    n = bodies.shape[0]
    src = gen.round_to_float(bodies) if prec == 32 else bodies
    lo, hi = ctx.shard_range()
    ctx.upload(src)
    f = np.zeros((n, dim))
    ctx.forces(out=f)
    ctx.step(dt, 1)
    after = src.copy()
    ctx.download(after)
    if world > 1:
        f = D.assemble_rows(f, lo, hi)
        after = D.assemble_rows(after, lo, hi)
    out = None
    if rank == 0:
        idx = np.sort(np.random.default_rng(2024).choice(n, min(PARITY_TARGETS, n), replace=False))
        ref = oracle.forces_targets(src, idx)
        kappa = oracle.condition_targets(src, idx) if prec == 32 else np.ones(idx.size)
        m = src[idx, 2 * dim:2 * dim + 1]
        implied = (after[idx, dim:2 * dim] - src[idx, dim:2 * dim]) * m / dt
        e_f = gen.relative_norm_error(f[idx], ref)
        e_s = gen.relative_norm_error(implied, ref)
        # the step's forces are read back through v1 - v0: allow for that cancellation in FP64
        cancel = np.abs(src[idx, dim:2 * dim]).max(axis=1) / np.maximum(np.abs(after[idx, dim:2 * dim] - src[idx, dim:2 * dim]).max(axis=1), 1e-300)
        e_s_adj = np.maximum(e_s - 4.5e-16 * cancel, 0.0)
        # x1 = x0 + v1 dt to the last bit or two (the device fuses the multiply-add), per body relative to its own scale
        step_dx = after[:, dim:2 * dim] * dt
        x_scale = np.maximum(np.maximum(np.abs(after[:, :dim]), np.abs(src[:, :dim])), np.abs(step_dx))
        x_res = (np.abs(after[:, :dim] - (src[:, :dim] + step_dx)) / np.maximum(x_scale, 1e-300)).max()
        ok = band_ok(e_f, kappa, prec) and band_ok(e_s_adj, kappa, prec) and x_res <= 4.5e-16
        out = {"targets": int(idx.size), "forces_max_rel_err": float(e_f.max()), "forces_p99": float(np.percentile(e_f, 99)),
               "step_max_rel_err": float(e_s_adj.max()), "step_p99": float(np.percentile(e_s_adj, 99)),
               "position_update_residual": float(x_res),
               "tol": ("max <= 1e-12 vs the CPU oracle" if prec == 64 else
                       "per body <= max(1e-5, 6e-7*kappa) and p99 <= 1e-5 vs the CPU oracle on the float-quantised inputs"),
               "kappa_max": float(kappa.max()) if prec == 32 else None}
        if world > 1:
            with pkg.NBodyCuda(dim, n, prec) as one:
                one.upload(src)
                f1 = one.forces()
                one.step(dt, 1)
                a1 = src.copy()
                one.download(a1)
            scale = np.abs(a1[:, :2 * dim]).max()
            dx = float(np.abs(after[:, :2 * dim] - a1[:, :2 * dim]).max() / scale)
            df = gen.relative_norm_error(f, f1)
            tol_x = 1e-12 if prec == 64 else 1e-6
            out["vs_1gpu"] = {"bodies": n, "state_max_diff_rel": dx, "forces_max_rel_diff": float(df.max()),
                              "forces_p99_rel_diff": float(np.percentile(df, 99)), "tol_state": tol_x}
            # forces: N GPUs evaluate the ordered pass per shard, one GPU the pair-symmetric pass -- two summation orders,
            # each within 1e-12 of the reference (checked above), hence within 2e-12 of each other
            ok = ok and dx <= tol_x and (df.max() <= 2e-12 if prec == 64 else np.percentile(df, 99) <= 1e-5)
            out["vs_1gpu"]["tol_forces"] = "max <= 2e-12" if prec == 64 else "p99 <= 1e-5"
        out["ok"] = bool(ok)
    ok_all = bcast_ok(out["ok"] if out else True)
    return out, ok_all


def energy_of(ctx):
    ke, pe = ctx.energy()
    return ke + pe


def run_config(pkg, oracle, name, prec_list=(32, 64)):
    """One BASELINE config on ONE GPU: throughput of a single nb200_step(nsteps) call, sampled-target parity
    at the initial positions, relative energy drift over the run for FP32 and FP64."""
    cfg = CONFIGS[name]
    gen = pkg.generators
    dim, n = cfg["dim"], cfg["n"]
    bodies = make_bodies(gen, cfg)
    steps_dist = cfg.get("steps_dist", cfg["dist"])
    steps_bodies = bodies if steps_dist == cfg["dist"] else make_bodies(gen, cfg, steps_dist)
    idx = np.sort(np.random.default_rng(7).choice(n, min(PARITY_TARGETS, n), replace=False))
    res = {"workload": f"{name}: brute-force {dim}D N={n} {cfg['dist']}, {cfg['steps']} steps, dt={cfg['dt']}"}
    if steps_dist != cfg["dist"]:
        res["steps_inputs"] = f"{steps_dist} (force parity on the {cfg['dist']} inputs; see CONFIGS in bench.py)"
    for prec in prec_list:
        src = gen.round_to_float(bodies) if prec == 32 else bodies
        with pkg.NBodyCuda(dim, n, prec) as ctx:
            ctx.upload(src)
            f = ctx.forces()
            ref = oracle.forces_targets(src, idx)
            err = gen.relative_norm_error(f[idx], ref)
            kappa = oracle.condition_targets(src, idx) if prec == 32 else np.ones(idx.size)
            steps_src = gen.round_to_float(steps_bodies) if prec == 32 else steps_bodies
            # Warm-up on a throw-away trajectory: 3 steps, then about 40 ms more of the same step so the SM
            # clocks are up after the CPU-side parity check (a 0.1 ms step timed from an idle GPU reads 6 %
            # slow); the inputs are uploaded again and ALL of the config's steps are timed in one call.
            ctx.upload(steps_src)
            ctx.step(cfg["dt"], 3)
            ctx.step(cfg["dt"], int(min(2000, max(1, 40.0 / max(ctx.last_elapsed_ms / 3, 1e-3)))))
            ctx.upload(steps_src)
            e0 = energy_of(ctx)
            ctx.step(cfg["dt"], cfg["steps"])
            ms = ctx.last_elapsed_ms / cfg["steps"]
            e1 = energy_of(ctx)
            res[f"f{prec}"] = {
                "value": round(n * (n - 1.0) / ms / 1e6, 1), "unit": "G interactions/s", "ms_per_step": round(ms, 4),
                "parity": {"targets": int(idx.size), "max_rel_err": float(err.max()), "p99": float(np.percentile(err, 99)),
                           "ok": band_ok(err, kappa, prec)},
                "energy_drift": float((e1 - e0) / e0), "plan": ctx.plan.split(" cutoff=")[0]}
    if "f32" in res and "f64" in res:
        d32, d64 = res["f32"]["energy_drift"], res["f64"]["energy_drift"]
        res["energy_drift_fp32_minus_fp64"] = d32 - d64
        res["ok"] = bool(res["f32"]["parity"]["ok"] and res["f64"]["parity"]["ok"] and
                         abs(d32 - d64) <= 1e-5 + 1e-2 * abs(d64))
    return res


def fp32_population_error(pkg, bodies, dim):
    """FP32-mode forces of ALL bodies against FP64 contexts on the device (nb200_compare_forces): (a) the same
    float-quantised inputs -- the arithmetic error the FP32 contract is stated on -- and (b) the unrounded
    inputs -- what the 24-bit quantisation of the positions adds -- and (c) the unrounded inputs again with option
    fp32_positions = 48, which removes that part (with the throughput of its forces call)."""
    gen = pkg.generators
    n = bodies.shape[0]
    rb = gen.round_to_float(bodies)
    out = {}
    keep = ("bodies", "max", "over_1e-5", "over_1e-4", "over_1e-3", "nonfinite", "histogram")
    with pkg.NBodyCuda(dim, n, pkg.NB200_FP32) as c32, pkg.NBodyCuda(dim, n, pkg.NB200_FP32) as c48:
        c32.upload(bodies)
        c32.forces()
        c48.set_option("fp32_positions", 48)        # (c) FP32 pair arithmetic on 48-bit positions (hi + lo float pairs)
        c48.upload(bodies)
        c48.forces()
        t0 = c48.last_elapsed_ms
        for key, src in (("vs_fp64_same_quantised_inputs", rb), ("vs_fp64_unrounded_inputs", bodies)):
            with pkg.NBodyCuda(dim, n, pkg.NB200_FP64) as c64:
                c64.upload(src)
                c64.forces()
                st = c32.compare_forces(c64)
                out[key] = {k: st[k] for k in keep}
                if src is bodies:
                    st = c48.compare_forces(c64)
                    out["fp32_positions_48_vs_fp64_unrounded_inputs"] = dict(
                        {k: st[k] for k in keep}, G_interactions_per_s=round(n * (n - 1.0) / t0 / 1e6, 1))
    return out


# ------------------------------------------------------------------------------------- product arm
def run_product_arm(args) -> None:
    import torch
    import torch.distributed as dist

    pkg = entry.load_package()
    from importlib import import_module
    D = import_module(pkg.__name__ + ".distributed")

    rank, world, local = D.env_rank_world()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cfg = dict(CONFIGS[args.config])
    if args.n:
        cfg["n"] = args.n
    n, prec, DIM, DT = cfg["n"], args.precision, cfg["dim"], cfg["dt"]
    bodies = make_bodies(pkg.generators, cfg)
    oracle = entry.load_oracle() if (rank == 0 and not args.no_parity) else None    # checker only, after the timed region
    if oracle is not None and world > 1:
        oracle.set_num_threads(max(1, (os.cpu_count() or 1) // 2))   # torchrun pinned OMP_NUM_THREADS=1; only rank 0 runs the checker

    def bcast_ok(flag: bool) -> bool:
        if world == 1:
            return flag
        t = torch.tensor([1 if flag else 0], device="cuda")
        dist.broadcast(t, src=0)
        return bool(int(t.item()))

    if world > 1:
        ctx = D.create_rank_context(pkg, DIM, n, prec, device=local)
    else:
        ctx = pkg.NBodyCuda(DIM, n, prec)
    for kv in args.opt or []:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    ctx.upload(bodies)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")    # > 126 MB L2

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up steps, then EXACTLY K steps, L2 flushed between them
    for _ in range(args.warmup):
        ctx.step(DT, 1)
    barrier()
    l0 = ctx.launch_count
    step_ms = []
    sampler = ClockSampler(local)
    with sampler:
        barrier()
        wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush_l2()
            ctx.step(DT, 1)
            step_ms.append(ctx.last_elapsed_ms)
        barrier()
        wall = time.perf_counter() - wall0
    launches = ctx.launch_count - l0
    dev_ms = max_over_ranks(sum(step_ms))
    value = interactions(n) * args.steps / (dev_ms * 1e-3) / 1e9
    clocks = sampler.summary()
    plan = ctx.plan

    # ---- the same K steps as ONE pipelined call (all-gather of step k+1 hidden behind pass A)
    barrier()
    ctx.step(DT, args.steps)
    pipe_ms = max_over_ranks(ctx.last_elapsed_ms)
    pipelined = interactions(n) * args.steps / (pipe_ms * 1e-3) / 1e9

    # ---- end to end through the host-buffer API: pinned host -> device, step, device -> host.
    # One host Body<D> array like the reference's std::vector<Body<D>>; under torchrun it lives in
    # shared memory (/dev/shm) mapped and page-locked by every rank, so each rank uploads all n
    # bodies, steps, and downloads the rows it owns straight into the common array.
    host_np, release_host = shared_pinned_bodies(bodies, rank, world, torch, barrier)
    lo, hi = ctx.shard_range()
    e2e_steps = max(1, min(args.steps, 3))
    ctx.upload(host_np)                     # untimed: first touch of the mapping
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.upload(host_np)
        ctx.step(DT, 1)
        ctx.download(host_np)
        barrier()                           # every rank's rows are in the common array
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_val = interactions(n) * e2e_steps / e2e_s / 1e9
    # N > 1: every rank copies only the rows it owns in both directions (shard-local upload, the default with the
    # peer-store exchange); the per-step figure is the sum over the ranks = one pass over the n bodies each way
    h2d = n * (2 * DIM + 1) * 8
    d2h = n * (2 * DIM + 1) * 8
    release_host()

    # ---- parity gate on the context that was just timed (every rank takes part; the oracle runs on rank 0)
    parity, parity_ok = (None, True)
    if not args.no_parity:
        parity, parity_ok = parity_gate(pkg, oracle, D, ctx, bodies, DIM, prec, DT, rank, world, local, bcast_ok)

    # ---- FP64 flavour of the same step (the reference's own precision), short
    fp64 = None
    if prec == 32 and not args.no_fp64:
        ctx.close()
        ctx = D.create_rank_context(pkg, DIM, n, pkg.NB200_FP64, device=local) if world > 1 else pkg.NBodyCuda(DIM, n, pkg.NB200_FP64)
        ctx.upload(bodies)
        ctx.step(DT, 1)
        barrier()
        ctx.step(DT, 2)
        ms64 = max_over_ranks(ctx.last_elapsed_ms)
        v64 = interactions(n) * 2 / (ms64 * 1e-3) / 1e9
        fp64 = {"value": round(v64, 2), "unit": "G interactions/s", "ms_per_step": round(ms64 / 2, 3), "steps": 2,
                "roofline_frac_fp64": round(v64 * 1e9 * FLOPS_PER_INTERACTION / (148 * 64 * 2 * 1.965e9 * world), 4)}
        if not args.no_parity:
            p64, ok64 = parity_gate(pkg, oracle, D, ctx, bodies, DIM, pkg.NB200_FP64, DT, rank, world, local, bcast_ok)
            parity_ok = parity_ok and ok64
            if p64 is not None:
                fp64["parity"] = p64
    ctx.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if not parity_ok:
            sys.exit(3)
        return

    # ---- the other BASELINE configs and the full-population FP32 error, 1 GPU only (short)
    sub_results, population = None, None
    if world == 1 and args.config == "c5" and not args.n and not args.no_sub:
        sub_results = {name: run_config(pkg, oracle or entry.load_oracle(), name) for name in ("c2", "c3", "c4")}
        if prec == 32:
            population = fp32_population_error(pkg, bodies, DIM)

    props = torch.cuda.get_device_properties(local)
    sms = props.multi_processor_count
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0)
    measured_peak = pkg.measure_fp32_peak(local) * world if prec == 32 else None   # live, CUDA events, this GPU
    traffic, traffic_src = ncu_traffic(n, prec) if world == 1 else (None, None)
    lanes = SM_LANES_FP32 if prec == 32 else 64
    peak_tflops = sms * lanes * 2 * sm_max * 1e6 / 1e12 * world
    achieved_tflops = value * 1e9 * FLOPS_PER_INTERACTION / 1e12
    roofline = {
        "bound": "fp32_fma_pipe" if prec == 32 else "fp64_pipe",
        "achieved": round(achieved_tflops, 2), "peak": round(peak_tflops, 2), "unit": "TFLOP/s",
        "frac": round(achieved_tflops / peak_tflops, 4),
        "peak_source": f"{sms} SMs x {lanes} FMA lanes x 2 x {sm_max:.0f} MHz x {world} GPU(s); MEASURED_PEAKS.json has only "
                       "HBM and bf16-tensor peaks, neither bounds this kernel (20 flops/interaction convention)",
        "frac_at_observed_clock": (round(achieved_tflops / (peak_tflops * clocks["sm_mhz"] / sm_max), 4)
                                   if clocks.get("sm_mhz") else None),
        "peak_measured": round(measured_peak, 2) if measured_peak else None,
        "frac_of_measured": round(achieved_tflops / measured_peak, 4) if measured_peak else None,
        "peak_measured_source": "nb200_measure_fp32_peak: independent packed FFMA2 chains timed with CUDA events in this run",
        # ordered pass: 11 FMA-pipe lane-ops per interaction; pair-symmetric pass: 15 per pair of interactions
        "fma_pipe_lane_ops_per_interaction": 7.5 if "pair-symmetric" in plan else 11,
        # what the pipe really did: executed lane-ops per interaction (FP32: 15 packed instructions per two pairs + the
        # hand-over adds of the rotation = 7.7; FP64: 18 DP instructions per pair = 9) x interactions / pipe capacity --
        # the figure ncu reports as pipe-active (profiles/), unlike `frac` not inflated by the 20-flop convention
        "pipe_utilisation": round(value * 1e9 * ((7.7 if "pair-symmetric" in plan else 11) if prec == 32 else
                                                 (9.0 if "pair-symmetric" in plan else 14.0))
                                  / (sms * lanes * sm_max * 1e6 * world), 4),
        "traffic": traffic, "traffic_source": traffic_src,
        "hbm_algorithmic_bytes_per_step": n * (16 if prec == 32 else 32) + n * (2 * DIM * 8 * 2 + 8 + 3 * 8 * 2),
    }
    line = {
        "metric": f"G body-body interactions/s (brute-force step, {DIM}D)", "value": round(value, 2),
        "unit": "G interactions/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dev_ms / args.steps, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32" if prec == 32 else "f64", "data": "synthetic",
        "config": {"workload": f"{args.config}: brute-force {DIM}D {cfg['dist']}, N={n}, fused force+integrate step, dt={DT}",
                   "n": n, "dim": DIM, "precision": prec,
                   "parallelism": (f"targets sharded over {world} GPU(s); each block of pairs evaluated by one of its two "
                                   "ranks, reaction sums and new positions stored into the peers' buffers over NVLink"
                                   if world > 1 else "1 GPU"),
                   "collective": ("none on the data path: peer stores over CUDA-IPC-mapped NVLink memory issued by the force/finish "
                                  "kernels + flag words; torch.distributed (NCCL) only bootstraps, barriers and reduces the timing"
                                  if world > 1 else "none"),
                   "l2": "flushed between timed steps (256 MiB write)", "plan": plan},
        "pipelined": {"value": round(pipelined, 2), "ms_per_step": round(pipe_ms / args.steps, 3),
                      "note": "same K steps in one nb200_step call, no L2 flush"},
        "wall_s_timed_region": round(wall, 3),
        "clocks": clocks,
        "e2e": {"value": round(e2e_val, 2), "unit": "G interactions/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "per_rank_bytes_each_way": (hi - lo) * (2 * DIM + 1) * 8},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if parity is not None:
        line["parity"] = parity
    if fp64:
        line["fp64"] = fp64
    if sub_results:
        line["sub_results"] = sub_results
    if population:
        line["fp32_error_all_bodies"] = population
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_reference_leg(min(n, CPU_SAMPLE_N), DIM, cfg)
    if world > 1:
        dist.destroy_process_group()
    print(json.dumps(line), flush=True)
    if not parity_ok or (sub_results and not all(r.get("ok", True) for r in sub_results.values())):
        print("bench.py: PARITY GATE FAILED (see the 'parity' / 'sub_results' objects of the line above)", file=sys.stderr)
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nb200", choices=["nb200", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS), help="BASELINE.json config timed as the headline (default c5)")
    ap.add_argument("--n", type=int, default=0, help="override the config's body count")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate after the timed region")
    ap.add_argument("--no-sub", action="store_true", help="skip the c2/c3/c4 sub-results and the full-population FP32 error")
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64])
    ap.add_argument("--opt", action="append", help="libnb200 option key=value (e.g. variant=1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fp64", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "nb200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_product_arm(args)


if __name__ == "__main__":
    main()
