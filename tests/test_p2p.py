"""The leaf (P2P) step of the suite's tree codes: nb200_p2p_leaves against
(a) golden vectors produced by the reference's own compiled BVH<D>::calculate_force on a single-leaf tree
    (tests/golden/p2p, oracle/gen_golden.py p2p; bvh.cpp:149-177),
(b) the oracle's restatement of that loop (oracle.p2p_leaves) on multi-leaf lists: a uniform cell grid whose cells are
    the leaves and whose 3^D neighbourhoods are the source lists -- the shape a tree code's near field takes.
Tolerance: FP64, per-body norm-wise relative error <= 1e-12 (the brute-force path's)."""
import glob
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P2P_GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "p2p", "*.npz")))
TOL = 1e-12


def single_leaf(n):
    return np.array([0, n]), np.arange(n), np.array([0, 1]), np.array([0])


def grid_leaves(bodies, dim, cells, rng=None, drop=0):
    """Leaves = the cells of a cells^dim grid over the bounding box (empty cells stay as empty leaves); source list of
    a leaf = the 3^dim cells around it, itself included.  `drop` bodies are left out of every leaf."""
    n = bodies.shape[0]
    x = bodies[:, :dim]
    lo, hi = x.min(axis=0), x.max(axis=0)
    c = np.minimum(((x - lo) / (hi - lo + 1e-300) * cells).astype(np.int64), cells - 1)
    key = np.zeros(n, dtype=np.int64)
    for d in range(dim):
        key = key * cells + c[:, d]
    members = np.arange(n)
    if drop:
        members = np.sort((rng or np.random.default_rng(0)).choice(n, n - drop, replace=False))
    order = members[np.argsort(key[members], kind="stable")]
    n_leaves = cells ** dim
    counts = np.bincount(key[order], minlength=n_leaves)
    leaf_off = np.concatenate([[0], np.cumsum(counts)])
    nbr_off, nbr = [0], []
    for leaf in range(n_leaves):
        idx = np.unravel_index(leaf, (cells,) * dim)
        for off in np.ndindex(*(3,) * dim):
            j = tuple(i + o - 1 for i, o in zip(idx, off))
            if all(0 <= v < cells for v in j):
                nbr.append(int(np.ravel_multi_index(j, (cells,) * dim)))
        nbr_off.append(len(nbr))
    return leaf_off, order, np.array(nbr_off), np.array(nbr)


# ------------------------------------------------------------------ CPU: the oracle is pinned
@pytest.mark.parametrize("path", P2P_GOLDEN, ids=[os.path.basename(p)[:-4] for p in P2P_GOLDEN])
def test_oracle_p2p_is_bit_identical_to_the_reference_bvh_leaf_loop(oracle, path):
    g = np.load(path)
    b = g["bodies"]
    f = oracle.p2p_leaves(b, *single_leaf(b.shape[0]), G=float(g["G"]), cutoff=float(g["cutoff"]), eps_same=float(g["eps_same"]))
    assert np.array_equal(f, g["forces_bvh_leaf"])
    if oracle.have_ref():
        assert np.array_equal(oracle.ref_bvh_single_leaf_forces(b), g["forces_bvh_leaf"])


def test_oracle_p2p_leaf_lists_add_up_to_the_direct_sum(oracle, pkg):
    """Size-independent property: with every cell in every source list the leaf sums are the all-pairs sum, i.e. minus
    the brute-force forces (the tree codes attract, the brute-force methods repel: SURVEY F2)."""
    b = pkg.generators.uniform_cube(600, 3, seed=5)
    leaf_off, order, _, _ = grid_leaves(b, 3, 3)
    n_leaves = leaf_off.shape[0] - 1
    nbr_off = np.arange(n_leaves + 1) * n_leaves
    nbr = np.tile(np.arange(n_leaves), n_leaves)
    f = oracle.p2p_leaves(b, leaf_off, order, nbr_off, nbr, cutoff=1e-10, eps_same=-1.0, skip_same_index=1)
    assert pkg.generators.relative_norm_error(f, -oracle.forces(b)).max() <= 1e-13
    # FMM flavour == BVH flavour away from the guards; sign = -1 is the brute-force convention
    f2 = oracle.p2p_leaves(b, leaf_off, order, nbr_off, nbr, cutoff=1e-10, eps_same=-1.0, skip_same_index=1, sign=-1)
    assert np.array_equal(f2, -f)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("path", P2P_GOLDEN, ids=[os.path.basename(p)[:-4] for p in P2P_GOLDEN])
def test_p2p_single_leaf_vs_reference_golden(pkg, path):
    g = np.load(path)
    b = g["bodies"]
    f = pkg.p2p_leaves_cuda(b, *single_leaf(b.shape[0]), G=float(g["G"]), **pkg.P2P_BVH)
    assert np.all(np.isfinite(f))
    e = pkg.generators.relative_norm_error(f, g["forces_bvh_leaf"])
    assert e.max() <= TOL, f"worst body {e.argmax()}: {e.max():.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("dim,n,cells,drop", [(3, 4000, 5, 0), (2, 3000, 7, 0), (3, 20000, 9, 37), (2, 5000, 2, 5), (3, 900, 1, 0)])
@pytest.mark.parametrize("flavour", ["bvh", "fmm"])
def test_p2p_leaf_lists_vs_oracle(pkg, oracle, dim, n, cells, drop, flavour):
    """Multi-leaf near-field lists (grid cells + 3^D neighbourhoods): ragged and empty leaves, leaves of more than 32
    bodies, bodies in no leaf, duplicates and pairs under the guards, both guard flavours and both signs."""
    rng = np.random.default_rng(n)
    b = pkg.generators.uniform_cube(n, dim, seed=n + dim)
    b[5, :dim] = b[4, :dim]
    b[9, :dim] = b[8, :dim]
    b[9, 0] += 2.0e-5
    b[13, :dim] = b[12, :dim]
    b[13, 0] += 4.0e-5
    lists = grid_leaves(b, dim, cells, rng, drop)
    params = pkg.P2P_BVH if flavour == "bvh" else pkg.P2P_FMM
    for sign in (1, -1):
        ref = oracle.p2p_leaves(b, *lists, sign=sign, **params)
        f, ms = pkg.p2p_leaves_cuda(b, *lists, sign=sign, return_ms=True, **params)
        assert np.all(np.isfinite(f)) and ms > 0.0
        in_leaf = np.zeros(n, dtype=bool)
        in_leaf[lists[1]] = True
        assert np.array_equal(f[~in_leaf], np.zeros_like(f[~in_leaf]))        # bodies in no leaf: Vector<D>()
        e = pkg.generators.relative_norm_error(f[in_leaf], ref[in_leaf])
        assert e.max() <= TOL, f"{flavour} sign={sign}: {e.max():.3e}"


@pytest.mark.gpu
def test_p2p_rejects_broken_lists(pkg):
    b = pkg.generators.uniform_cube(100, 3, seed=1)
    lo, lb, no, nl = single_leaf(100)
    for bad in ((lo, np.full(100, 100), no, nl),            # body index out of range
                (lo, np.zeros(100, dtype=np.int64), no, nl),     # a body in two leaf slots
                (np.array([0, 50, 40]), lb, np.array([0, 1, 2]), np.array([0, 1])),   # decreasing offsets
                (lo, lb, no, np.array([3]))):                 # neighbour leaf out of range
        with pytest.raises(pkg.NB200Error):
            pkg.p2p_leaves_cuda(b, *bad)
    assert pkg.p2p_leaves_cuda(b, np.array([0]), np.array([], dtype=np.int64), np.array([0]), np.array([], dtype=np.int64)).shape == (100, 3)
