#!/usr/bin/env python
"""Print the CUDA-event timeline of a pipelined multi-GPU nb200_step call (single process, --ngpus G)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--ngpus", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--precision", type=int, default=32)
ap.add_argument("--overlap", type=int, default=1)
a = ap.parse_args()
pkg = entry.load_package()
b = pkg.generators.uniform_cube(a.n, 3, seed=1)
with pkg.NBodyCuda(3, a.n, a.precision, a.ngpus) as ctx:
    ctx.set_option("overlap", a.overlap)
    ctx.upload(b)
    ctx.step(1e-3, 2)
    ctx.set_option("trace", 1)
    ctx.step(1e-3, a.steps)
    print(f"ngpus={a.ngpus} overlap={a.overlap} ms/step={ctx.last_elapsed_ms / a.steps:.3f}")
    print(ctx.plan)
