#!/usr/bin/env python
"""Run the BASELINE.json configurations end to end and print one JSON report line per config.

    python tools/config_run.py --config c2|c3|c4|c5|all [--ngpus G] [--precision 32|64] [--steps K] [--no-cpu]

  c2  3D  N=16384   uniform cube, 100 steps, 1 GPU, next to the reference's OpenMP / ParlayLib paths
  c3  2D  N=65536   uniform (stratified) square, 100 steps, 1 GPU (the 2D Vector specialisation)
  c4  3D  N=262144  Plummer sphere, 100 steps, G GPUs, energy-drift check
  c5  3D  N=2^20    uniform cube, 10 steps, G GPUs (the headline; bench.py measures it properly)

For every config: device-timed throughput of ONE nb200_step(nsteps) call, relative energy drift
(E(t)-E(0))/E(0) of the run, and force parity of 256 sampled targets against the CPU oracle
(all sources) before the first and after the last step.  FP32 runs are also compared with an
FP64 run of the same configuration (drift and positions): at these sizes a CPU trajectory would
take hours, so FP64-GPU -- itself pinned to the oracle -- is the trajectory reference.
Multi-GPU here is the single-process flavour (nb200_create with ngpus=G, fused NVLink exchange).
The oracle is used as the CHECKER only.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

CONFIGS = {
    # name: (dim, n, distribution, steps, dt)
    "c2": (3, 16384, "cube", 100, 1e-4),
    "c3": (2, 65536, "jittered", 100, 1e-5),
    "c4": (3, 262144, "plummer", 100, 1e-3),
    "c5": (3, 1 << 20, "cube", 10, 1e-6),
}


def make_bodies(gen, dim, n, dist, seed):
    if dist == "plummer":
        return gen.plummer(n, seed=seed)
    if dist == "jittered":      # stratified uniform: Poisson-uniform 2D points hold pairs no fixed dt resolves
        return gen.jittered_cube(n, dim, seed=seed)
    return gen.uniform_cube(n, dim, seed=seed)


def sampled_parity(pkg, oracle, bodies, forces, prec, k=256):
    n = bodies.shape[0]
    idx = np.random.default_rng(0).choice(n, min(k, n), replace=False)
    src = pkg.generators.round_to_float(bodies) if prec == 32 else bodies
    ref = oracle.forces_targets(src, idx)
    err = pkg.generators.relative_norm_error(forces[idx], ref)
    return {"targets": int(idx.size), "max": float(err.max()), "p50": float(np.median(err)),
            "p99": float(np.percentile(err, 99))}


def run_one(pkg, oracle, name, prec, ngpus, steps_override=None, chunks=5, dt_override=0.0):
    dim, n, dist, steps, dt = CONFIGS[name]
    dt = dt_override or dt
    if steps_override:
        steps = steps_override
    gen = pkg.generators
    bodies = make_bodies(gen, dim, n, dist, seed=42 + int(name[1]))
    inter = float(n) * (n - 1)
    out = {"config": name, "dim": dim, "n": n, "distribution": dist, "steps": steps, "dt": dt,
           "precision": prec, "ngpus": ngpus}
    with pkg.NBodyCuda(dim, n, prec, ngpus) as ctx:
        ctx.upload(bodies)
        f0 = ctx.forces()
        out["force_parity_step0"] = sampled_parity(pkg, oracle, bodies, f0, prec)
        ke0, pe0 = ctx.energy()
        e0 = ke0 + pe0
        drift, ms = [], 0.0
        per = max(1, steps // chunks)
        done = 0
        ctx.step(dt, 0)
        while done < steps:
            k = min(per, steps - done)
            ctx.step(dt, k)
            ms += ctx.last_elapsed_ms
            done += k
            drift.append((sum(ctx.energy()) - e0) / e0)
        out["plan"] = ctx.plan
        out["ms_per_step"] = round(ms / steps, 4)
        out["G_interactions_per_s"] = round(inter * steps / (ms * 1e-3) / 1e9, 1)
        out["energy0"] = {"kinetic": ke0, "potential": pe0}
        out["energy_drift"] = [float(d) for d in drift]
        after = bodies.copy()
        ctx.download(after)
        ctx.upload(after)
        f1 = ctx.forces()
        out["force_parity_last_step"] = sampled_parity(pkg, oracle, after, f1, prec)
    return out, after


def cpu_leg(oracle, pkg, name):
    """The reference's own parallel brute-force variants on this host: one force evaluation each."""
    dim, n, dist, _, _ = CONFIGS[name]
    if not oracle.have_ref():
        return {"unavailable": "oracle/_ref/libnbref.so not present"}
    bodies = make_bodies(pkg.generators, dim, n, dist, seed=42 + int(name[1]))
    thr = oracle.ref_threads()
    res = {"threads": thr, "cores": os.cpu_count()}
    for v in ("omp_1", "omp_2", "parlay_1", "parlay_2"):
        t = oracle.ref_forces(bodies, v, want_forces=False)[1]
        res[v] = {"seconds": round(t, 4), "G_interactions_per_s": round(float(n) * (n - 1) / t / 1e9, 3)}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="all")
    ap.add_argument("--ngpus", type=int, default=1)
    ap.add_argument("--precision", type=int, default=32)
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--dt", type=float, default=0.0, help="override the config's time step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fp64-ref", action="store_true")
    a = ap.parse_args()
    pkg = entry.load_package()
    oracle = entry.load_oracle()
    names = list(CONFIGS) if a.config == "all" else a.config.split(",")
    for name in names:
        t0 = time.time()
        chunks = 2 if name == "c5" else 5
        rep, after = run_one(pkg, oracle, name, a.precision, a.ngpus, a.steps or None, chunks, a.dt)
        if a.precision == 32 and not a.no_fp64_ref:
            ref64, after64 = run_one(pkg, oracle, name, 64, a.ngpus, a.steps or None, chunks, a.dt)
            dim = rep["dim"]
            scale = np.abs(after64[:, :dim]).max()
            rep["vs_fp64_run"] = {
                "fp64_G_interactions_per_s": ref64["G_interactions_per_s"],
                "fp64_energy_drift": ref64["energy_drift"],
                "fp64_force_parity_step0": ref64["force_parity_step0"],
                "fp64_force_parity_last_step": ref64["force_parity_last_step"],
                "max_position_diff_rel": float(np.abs(after[:, :dim] - after64[:, :dim]).max() / scale),
                "drift_diff_last": float(rep["energy_drift"][-1] - ref64["energy_drift"][-1]),
            }
        if name in ("c2", "c3") and not a.no_cpu:
            rep["cpu_reference"] = cpu_leg(oracle, pkg, name)
        rep["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(rep), flush=True)


if __name__ == "__main__":
    main()
