#!/usr/bin/env python
"""Time the FP32 pair-symmetric kernel flavours (option sym_algo) of one libnb200 build and check each
against the oracle on sampled targets.  One JSON line per (case, algo).

    NB200_LIB=.../libnb200_x.so python tools/sym_variants.py [--cases 3:1048576,2:65536] [--algos 0,1,2]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="3:1048576,3:262144,2:65536")
    ap.add_argument("--algos", default="0,1,2")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--precision", type=int, default=32)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--per-call", type=int, default=1, help="steps per nb200_step call (the timed quantity is ms per step)")
    a = ap.parse_args()
    pkg = entry.load_package()
    oracle = entry.load_oracle()
    gen = pkg.generators
    lib = os.path.basename(os.environ.get("NB200_LIB", "libnb200.so"))
    for case in a.cases.split(","):
        dim, n = (int(x) for x in case.split(":"))
        bodies = gen.uniform_cube(n, dim, seed=47)
        src = gen.round_to_float(bodies) if a.precision == 32 else bodies
        idx = np.random.default_rng(1).choice(n, 512, replace=False)
        ref = oracle.forces_targets(src, idx)
        for algo in [int(x) for x in a.algos.split(",")]:
            with pkg.NBodyCuda(dim, n, a.precision) as ctx:
                ctx.set_option("sym_algo", algo)
                for kv in a.opt:
                    k, v = kv.split("=")
                    ctx.set_option(k, int(v))
                ctx.upload(src)
                f = ctx.forces()
                err = gen.relative_norm_error(f[idx], ref)
                ctx.step(1e-6, 1)
                ms = []
                for _ in range(a.steps):
                    ctx.step(1e-6, a.per_call)
                    ms.append(ctx.last_elapsed_ms / a.per_call)
                best = min(ms)
                print(json.dumps({"lib": lib, "dim": dim, "n": n, "algo": algo, "ms_per_step": round(best, 4),
                                  "G_inter_per_s": round(n * (n - 1.0) / best / 1e6, 1),
                                  "err_max": float(err.max()), "err_p99": float(np.percentile(err, 99)),
                                  "finite": bool(np.all(np.isfinite(f))), "plan": ctx.plan[:160]}), flush=True)


if __name__ == "__main__":
    main()
