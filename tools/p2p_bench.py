#!/usr/bin/env python
"""Throughput of the leaf (P2P) step on near-field lists of the shape a tree code produces: cells of a uniform grid
as leaves (~16 bodies each, the reference BVH's default leaf size), 3^D neighbourhoods as source lists; the oracle's
restatement of the reference's leaf loop on all host threads beside it.  One JSON line.
    python tools/p2p_bench.py [--n 1048576] [--dim 3] [--leaf 16]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_p2p import grid_leaves

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--dim", type=int, default=3)
ap.add_argument("--leaf", type=int, default=16)
ap.add_argument("--cpu-leaves", type=int, default=4096, help="leaves of the CPU sample")
a = ap.parse_args()
pkg = e.load_package(); oracle = e.load_oracle(); gen = pkg.generators
b = gen.uniform_cube(a.n, a.dim, seed=9)
cells = max(1, int(round((a.n / a.leaf) ** (1.0 / a.dim))))
lists = grid_leaves(b, a.dim, cells)
leaf_off, order, nbr_off, nbr = lists
sizes = np.diff(leaf_off)
pairs = float(sum(sizes[l] * sizes[nbr[nbr_off[l]:nbr_off[l + 1]]].sum() for l in range(sizes.shape[0])))
f, ms = pkg.p2p_leaves_cuda(b, *lists, return_ms=True, **pkg.P2P_BVH)     # warm-up (allocations, module load)
best = min(pkg.p2p_leaves_cuda(b, *lists, return_ms=True, **pkg.P2P_BVH)[1] for _ in range(3))
# CPU: the first --cpu-leaves target leaves (their source lists reach into the rest), all host threads
k = min(a.cpu_leaves, sizes.shape[0])
sub_nbr_off = np.concatenate([nbr_off[:k + 1], np.full(sizes.shape[0] - k, nbr_off[k])])     # later leaves: no sources
t0 = time.perf_counter()
ref = oracle.p2p_leaves(b, leaf_off, order, sub_nbr_off, nbr, **{"cutoff": 1e-9, "eps_same": 1e-9})
cpu_s = time.perf_counter() - t0
cpu_pairs = float(sum(sizes[l] * sizes[nbr[nbr_off[l]:nbr_off[l + 1]]].sum() for l in range(k)))
tg = order[:leaf_off[k]]
err = gen.relative_norm_error(f[tg], ref[tg]).max()
print(json.dumps({"n": a.n, "dim": a.dim, "leaves": int(sizes.shape[0]), "mean_leaf": float(sizes.mean()), "max_leaf": int(sizes.max()),
                  "pair_interactions": pairs, "gpu_kernel_ms": round(best, 3), "gpu_G_inter_per_s": round(pairs / best / 1e6, 1),
                  "cpu_sample_leaves": k, "cpu_G_inter_per_s": round(cpu_pairs / cpu_s / 1e9, 3), "cpu_threads": oracle.num_threads(),
                  "max_rel_err_vs_oracle_on_sample": float(err)}))
