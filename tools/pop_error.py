#!/usr/bin/env python
"""Full-population FP32 error of the pair-symmetric flavours against an FP64 context on the same quantised
inputs (nb200_compare_forces), plus the worst err/kappa ratio on a CPU-checkable case.
    python tools/pop_error.py [--n 1048576] [--dim 3]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--dim", type=int, default=3)
a = ap.parse_args()
pkg = e.load_package(); oracle = e.load_oracle(); gen = pkg.generators
cases = [{"symmetric": 0}, {"sym_algo": 0, "sym_ti": 4, "sym_block": 256}, {"sym_algo": 1, "sym_ti": 4, "sym_block": 256},
         {"sym_algo": 1, "sym_ti": 8}, {"sym_algo": 2, "sym_ti": 8}, {"sym_algo": 2, "sym_ti": 4, "sym_block": 256}]
b = gen.round_to_float(gen.uniform_cube(a.n, a.dim, seed=47))
with pkg.NBodyCuda(a.dim, a.n, 64) as c64:
    c64.upload(b); c64.forces()
    for opts in cases:
        with pkg.NBodyCuda(a.dim, a.n, 32) as c32:
            for k, v in opts.items():
                c32.set_option(k, v)
            c32.set_option("detect", 1)
            c32.upload(b); c32.forces()
            st = c32.compare_forces(c64)
        print(json.dumps({"n": a.n, "dim": a.dim, "opts": opts, "max": st["max"], "over_1e-5": st["over_1e-5"], "over_1e-4": st["over_1e-4"],
                          "hist": st["histogram"]}), flush=True)
# kappa-normalised worst case where the full oracle is affordable
for dim, n, seed in ((2, 20000, 44), (3, 16384, 44), (2, 65536, 45)):
    bb = gen.round_to_float(gen.uniform_cube(n, dim, seed=seed))
    ref = oracle.forces(bb); kappa = oracle.condition(bb)
    for opts in cases:
        f = pkg.brute_force_cuda_n_body(bb, 32, options=dict(opts, detect=1))
        err = gen.relative_norm_error(f, ref)
        print(json.dumps({"n": n, "dim": dim, "opts": opts, "max_err": float(err.max()), "max_err_over_kappa": float((err / kappa).max()),
                          "worst_vs_band": float((err / np.maximum(1e-5, 6e-7 * kappa)).max()), "p99": float(np.percentile(err, 99))}), flush=True)
