#!/usr/bin/env python
"""Run one small configuration of the hot path (for ncu captures, sweeps and quick timings).

    python tools/run_case.py --n 131072 --dim 3 --precision 32 --steps 2 [--opt variant=0 ...] [--dist cube|plummer]
Prints one JSON line with the device-timed throughput of the LAST nb200_step call.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=131072)
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--precision", type=int, default=32)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--ngpus", type=int, default=1)
    ap.add_argument("--dist", default="cube")
    ap.add_argument("--opt", action="append")
    ap.add_argument("--sweep", action="store_true", help="sweep variant x seg_tiles and print a table")
    a = ap.parse_args()
    pkg = entry.load_package()
    gen = pkg.generators
    bodies = gen.plummer(a.n, seed=1) if a.dist == "plummer" else gen.uniform_cube(a.n, a.dim, seed=1)
    inter = float(a.n) * (a.n - 1)

    def one(opts):
        with pkg.NBodyCuda(a.dim, a.n, a.precision, a.ngpus) as ctx:
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.upload(bodies)
            if a.warmup:
                ctx.step(1e-3, a.warmup)
            ctx.step(1e-3, a.steps)
            ms = ctx.last_elapsed_ms / a.steps
            return {"n": a.n, "dim": a.dim, "precision": a.precision, "ms_per_step": round(ms, 4),
                    "G_inter_per_s": round(inter / ms / 1e6, 1), "plan": ctx.plan}

    base = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in (a.opt or []))
    if not a.sweep:
        print(json.dumps(one(base)))
        return
    for variant in range(5):
        for seg in (0, 4, 16, 64):
            r = one(dict(base, variant=variant, seg_tiles=seg))
            print(f"variant={variant} seg={seg:3d}  {r['ms_per_step']:10.4f} ms  {r['G_inter_per_s']:9.1f} G/s   {r['plan']}")


if __name__ == "__main__":
    main()
