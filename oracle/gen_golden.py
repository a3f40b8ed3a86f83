#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's OWN compiled brute-force code.

Run in the build container (needs /root/reference):  python oracle/gen_golden.py
Every array below is produced by oracle/_ref/libnbref.so, i.e. by the unmodified
/root/reference/nbody-sim-new/methods.cpp (brute_force_seq_n_body :7-42,
brute_force_omp_n_body_1 :45-95, brute_force_omp_n_body_2 :98-136,
update_body_velocities :426-438, update_body_positions :441-450) with the reference's
compile-time G (utils.h:21) and 1e-10 pair cut-off.  The reference ships no golden vectors of
its own; these fixtures are what pins both oracle/nbody_oracle.c and the
CUDA path on machines where /root/reference does not exist (the GPU box).
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "nb_generators", os.path.join(ROOT, "nbody-simulation-parallel_b200", "generators.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

OUT = os.path.join(ROOT, "tests", "golden")


def degenerate(dim):
    """Hand-made edge cases for the hard cut-off (methods.cpp:119: dist_sq < 1e-10 -> skip):
    exact duplicates (r^2 = 0, i != j), a pair just inside (r^2 = 0.81e-10) and just outside
    (r^2 = 1.21e-10) the cut-off, a zero-velocity body, widely different masses."""
    pts = [
        [0.25, 0.25, 0.25], [0.25, 0.25, 0.25],               # duplicates -> skipped
        [0.75, 0.50, 0.50], [0.75 + 0.9e-5, 0.50, 0.50],      # r^2 = 0.81e-10 -> skipped
        [0.50, 0.75, 0.10], [0.50, 0.75 + 1.1e-5, 0.10],      # r^2 = 1.21e-10 -> kept (huge force)
        [0.10, 0.90, 0.80], [0.90, 0.10, 0.30], [0.40, 0.60, 0.95],
    ]
    n = len(pts)
    b = np.zeros((n, 2 * dim + 1))
    b[:, :dim] = np.array(pts)[:, :dim]
    rng = np.random.default_rng(7)
    b[:, dim:2 * dim] = rng.uniform(-0.1, 0.1, (n, dim))
    b[6, dim:2 * dim] = 0.0
    b[:, 2 * dim] = np.array([1.0, 2.0, 3.0, 0.5, 1e3, 1e-3, 1.0, 7.0, 0.25]) / (oracle.G_REF * n)
    return b


def main():
    os.makedirs(OUT, exist_ok=True)
    if not oracle.have_ref():
        oracle.build(ref=True)
    cases = {
        # BASELINE.json configs[0]: seq 3D N=1024, 10 steps, reference-range inputs
        "c1_refrange3d_n1024": (gen.reference_range(1024, 3, seed=43), 1.0, 10),
        "refrange2d_n300": (gen.reference_range(300, 2, seed=44), 1.0, 10),
        "cube3d_n1000": (gen.uniform_cube(1000, 3, seed=45), 1e-4, 20),   # ragged: not a tile multiple
        "cube2d_n1280": (gen.uniform_cube(1280, 2, seed=46), 1e-5, 20),
        "plummer3d_n768": (gen.plummer(768, seed=47), 1e-3, 20),
        "degenerate3d_n9": (degenerate(3), 1e-4, 5),
        "degenerate2d_n9": (degenerate(2), 1e-4, 5),
        "tiny3d_n1": (gen.uniform_cube(1, 3, seed=48), 1e-3, 3),
        "tiny2d_n2": (gen.uniform_cube(2, 2, seed=49), 1e-3, 3),
        "tiny3d_n3": (gen.uniform_cube(3, 3, seed=50), 1e-3, 3),
        "ragged3d_n257": (gen.uniform_cube(257, 3, seed=51), 1e-4, 5),
    }
    for name, (bodies, dt, nsteps) in cases.items():
        f_seq, _ = oracle.ref_forces(bodies, "seq")
        f_omp1, _ = oracle.ref_forces(bodies, "omp_1")
        f_omp2, _ = oracle.ref_forces(bodies, "omp_2")
        after_seq = oracle.ref_simulate(bodies, dt, nsteps, "seq")
        after_omp2 = oracle.ref_simulate(bodies, dt, nsteps, "omp_2")
        np.savez(os.path.join(OUT, name + ".npz"), bodies=bodies, forces_seq=f_seq,
                 forces_omp1=f_omp1, forces_omp2=f_omp2, after_seq=after_seq,
                 after_omp2=after_omp2, dt=np.float64(dt), nsteps=np.int64(nsteps),
                 G=np.float64(oracle.ref_threads()["G"]), cutoff=np.float64(1e-10))
        print(f"{name}: n={bodies.shape[0]} dim={(bodies.shape[1] - 1) // 2} "
              f"|F|max={np.abs(f_omp2).max():.3e}")


def p2p_cases():
    """Inputs for the leaf (P2P) step of the tree codes: the reference's ranges and the unit cube, each with an exact
    duplicate (the per-component same-position test, bvh.cpp:156-163), a pair under the 1e-9 guard on r^2
    (bvh.cpp:170) and a pair just above it."""
    out = {}
    for name, b in (("p2p_bvh3d_refrange_n300", gen.reference_range(300, 3, seed=61)),
                    ("p2p_bvh2d_refrange_n300", gen.reference_range(300, 2, seed=62)),
                    ("p2p_bvh3d_cube_n500", gen.uniform_cube(500, 3, seed=63)),
                    ("p2p_bvh2d_cube_n77", gen.uniform_cube(77, 2, seed=64))):
        dim = (b.shape[1] - 1) // 2
        b[5, :dim] = b[4, :dim]                       # same position: skipped
        b[9, 0] = b[8, 0] + 2.0e-5                    # r^2 = 4e-10 < 1e-9 ...
        b[9, 1:dim] = b[8, 1:dim]                     # ... skipped
        b[13, 0] = b[12, 0] + 4.0e-5                  # r^2 = 1.6e-9: kept (dominates both bodies)
        b[13, 1:dim] = b[12, 1:dim]
        out[name] = b
    return out


def p2p_main():
    """tests/golden/p2p/*.npz: BVH<D>::calculate_force of the compiled reference on a tree whose root is a single leaf
    (max_bodies_per_leaf >= n): its leaf loop (bvh.cpp:149-177) as the direct sum over all bodies."""
    out_dir = os.path.join(OUT, "p2p")
    os.makedirs(out_dir, exist_ok=True)
    for name, bodies in p2p_cases().items():
        f = oracle.ref_bvh_single_leaf_forces(bodies)
        np.savez(os.path.join(out_dir, name + ".npz"), bodies=bodies, forces_bvh_leaf=f,
                 G=np.float64(oracle.ref_threads()["G"]), cutoff=np.float64(1e-9), eps_same=np.float64(1e-9))
        print(f"{name}: n={bodies.shape[0]} |F|max={np.abs(f).max():.3e}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "p2p":
        p2p_main()          # the brute-force fixtures stay as committed
    else:
        main()
        p2p_main()
