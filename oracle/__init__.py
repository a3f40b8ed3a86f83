"""ctypes bindings for the CHECKERS under oracle/ (test infrastructure only).

* ``liboracle.so`` (oracle/_build) -- our plain-C restatement, oracle/nbody_oracle.c.
* ``libnbref.so``  (oracle/_ref)   -- the reference's own methods.cpp compiled unmodified
  (oracle/Makefile); present where it was built (this container) and on the GPU box only as
  the prebuilt file that travelled with the snapshot.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product (libnb200.so and the package
``nbody-simulation-parallel_b200``) never does.

Bodies are numpy float64 arrays of shape (n, 2*D+1): the reference's AoS ``Body<D>``
(body.h:7-19) -- position[D], velocity[D], mass.  Forces are (n, D) = ``Vector<D>``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libnbref.so")

#: utils.h:21 and methods.cpp:24 -- the reference's compile-time constants.
G_REF = 4.471e-21
CUTOFF_REF = 1e-10

VARIANTS = {"seq": 0, "omp_1": 1, "omp_2": 2, "parlay_1": 3, "parlay_2": 4}

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)


def build(ref: bool = True) -> None:
    """Compile the checkers (building the checker is not using it)."""
    target = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + target, check=True)


def _as_bodies(bodies, dim):
    b = np.ascontiguousarray(bodies, dtype=np.float64)
    if b.ndim != 2 or b.shape[1] != 2 * dim + 1:
        raise ValueError(f"bodies must be (n, {2 * dim + 1}) for dim={dim}, got {b.shape}")
    return b


def _dim_of(bodies):
    w = np.asarray(bodies).shape[1]
    if w not in (5, 7):
        raise ValueError("bodies must have 5 (2D) or 7 (3D) columns")
    return (w - 1) // 2


class _Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = L = ctypes.CDLL(ORACLE_SO)
        sz = ctypes.c_size_t
        for name in ("oracle_forces_seq", "oracle_forces_omp2"):
            f = getattr(L, name)
            f.argtypes = [ctypes.c_int, sz, _dp, ctypes.c_double, ctypes.c_double, _dp]
            f.restype = ctypes.c_int
        for name in ("oracle_forces_targets", "oracle_forces_targets_ld"):
            f = getattr(L, name)
            f.argtypes = [ctypes.c_int, sz, _dp, ctypes.c_double, ctypes.c_double, _ip, sz, _dp]
            f.restype = ctypes.c_int
        L.oracle_update_velocities.argtypes = [ctypes.c_int, sz, _dp, _dp, ctypes.c_double]
        L.oracle_update_positions.argtypes = [ctypes.c_int, sz, _dp, ctypes.c_double]
        L.oracle_simulate.argtypes = [ctypes.c_int, sz, _dp, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.oracle_energy.argtypes = [ctypes.c_int, sz, _dp, ctypes.c_double, ctypes.c_double, _dp, _dp]
        L.oracle_condition.argtypes = [ctypes.c_int, sz, _dp, ctypes.c_double, ctypes.c_double, _dp]
        L.oracle_condition_targets.argtypes = [ctypes.c_int, sz, _dp, ctypes.c_double, ctypes.c_double, _ip, sz, _dp]
        L.oracle_p2p_leaves.argtypes = [ctypes.c_int, sz, _dp, sz, _ip, _ip, _ip, _ip, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_int, ctypes.c_int, _dp]
        L.oracle_accuracy_pct.argtypes = [ctypes.c_int, sz, _dp, _dp]
        L.oracle_accuracy_pct.restype = ctypes.c_double
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        L.oracle_set_num_threads.restype = None


_oracle = None


def _o():
    global _oracle
    if _oracle is None:
        _oracle = _Oracle()
    return _oracle.lib


def _p(a):
    return a.ctypes.data_as(_dp)


def forces(bodies, G=G_REF, cutoff=CUTOFF_REF, variant="omp_2"):
    """Reference brute-force forces, restated: variant 'seq' (methods.cpp:7-42) or
    'omp_2' (methods.cpp:98-136, the canonical oracle)."""
    dim = _dim_of(bodies)
    b = _as_bodies(bodies, dim)
    out = np.zeros((b.shape[0], dim))
    fn = _o().oracle_forces_seq if variant == "seq" else _o().oracle_forces_omp2
    rc = fn(dim, b.shape[0], _p(b), G, cutoff, _p(out))
    if rc:
        raise RuntimeError(f"oracle forces rc={rc}")
    return out


def forces_targets(bodies, targets, G=G_REF, cutoff=CUTOFF_REF, long_double=False):
    """omp_2 row sums for a subset of targets (all sources)."""
    dim = _dim_of(bodies)
    b = _as_bodies(bodies, dim)
    t = np.ascontiguousarray(targets, dtype=np.int64)
    out = np.zeros((t.shape[0], dim))
    fn = _o().oracle_forces_targets_ld if long_double else _o().oracle_forces_targets
    rc = fn(dim, b.shape[0], _p(b), G, cutoff, t.ctypes.data_as(_ip), t.shape[0], _p(out))
    if rc:
        raise RuntimeError(f"oracle forces_targets rc={rc}")
    return out


def simulate(bodies, dt, nsteps, G=G_REF, cutoff=CUTOFF_REF, variant="omp_2"):
    """nsteps of {force; update_body_velocities; update_body_positions}.
    Returns a new (n, 2D+1) array."""
    dim = _dim_of(bodies)
    b = _as_bodies(bodies, dim).copy()
    rc = _o().oracle_simulate(dim, b.shape[0], _p(b), G, cutoff, dt, nsteps,
                              0 if variant == "seq" else 1)
    if rc:
        raise RuntimeError(f"oracle simulate rc={rc}")
    return b


def energy(bodies, G=G_REF, cutoff=CUTOFF_REF):
    """(kinetic, potential) of the reference's own r^-3 repulsive law: U = sum G m m / (2 r^2)."""
    dim = _dim_of(bodies)
    b = _as_bodies(bodies, dim)
    ke, pe = ctypes.c_double(), ctypes.c_double()
    rc = _o().oracle_energy(dim, b.shape[0], _p(b), G, cutoff, ctypes.byref(ke), ctypes.byref(pe))
    if rc:
        raise RuntimeError(f"oracle energy rc={rc}")
    return ke.value, pe.value


def condition(bodies, G=G_REF, cutoff=CUTOFF_REF):
    """kappa_i = sum_j |f_ij| / |sum_j f_ij|: the summation condition number of each body's force."""
    dim = _dim_of(bodies)
    b = _as_bodies(bodies, dim)
    out = np.ones(b.shape[0])
    rc = _o().oracle_condition(dim, b.shape[0], _p(b), G, cutoff, _p(out))
    if rc:
        raise RuntimeError(f"oracle condition rc={rc}")
    return out


def condition_targets(bodies, targets, G=G_REF, cutoff=CUTOFF_REF):
    """kappa of a subset of targets (all sources)."""
    dim = _dim_of(bodies)
    b = _as_bodies(bodies, dim)
    t = np.ascontiguousarray(targets, dtype=np.int64)
    out = np.ones(t.shape[0])
    rc = _o().oracle_condition_targets(dim, b.shape[0], _p(b), G, cutoff, t.ctypes.data_as(_ip), t.shape[0], _p(out))
    if rc:
        raise RuntimeError(f"oracle condition_targets rc={rc}")
    return out


def p2p_leaves(bodies, leaf_offsets, leaf_bodies, nbr_offsets, nbr_leaves, G=G_REF, cutoff=1e-9, eps_same=1e-9,
               skip_same_index=0, sign=1):
    """The leaf (P2P) branch of the reference's tree codes, restated (bvh.cpp:149-177; fmm.cpp:622-637 with
    eps_same=-1, cutoff=1e-10, skip_same_index=1): direct sums over leaf lists."""
    dim = _dim_of(bodies)
    b = _as_bodies(bodies, dim)
    lo = np.ascontiguousarray(leaf_offsets, dtype=np.int64)
    lb = np.ascontiguousarray(leaf_bodies, dtype=np.int64)
    no = np.ascontiguousarray(nbr_offsets, dtype=np.int64)
    nl = np.ascontiguousarray(nbr_leaves, dtype=np.int64)
    out = np.zeros((b.shape[0], dim))
    rc = _o().oracle_p2p_leaves(dim, b.shape[0], _p(b), lo.shape[0] - 1, lo.ctypes.data_as(_ip), lb.ctypes.data_as(_ip),
                                no.ctypes.data_as(_ip), nl.ctypes.data_as(_ip), G, cutoff, eps_same, int(skip_same_index),
                                int(sign), _p(out))
    if rc:
        raise RuntimeError(f"oracle p2p_leaves rc={rc}")
    return out


def accuracy_pct(forces_, reference):
    """utils.h:170-219 (the reference's -a 1 column)."""
    f = np.ascontiguousarray(forces_, dtype=np.float64)
    r = np.ascontiguousarray(reference, dtype=np.float64)
    return _o().oracle_accuracy_pct(f.shape[1], f.shape[0], _p(f), _p(r))


def num_threads():
    return _o().oracle_num_threads()


def set_num_threads(n: int):
    """Checker threads (torchrun pins OMP_NUM_THREADS=1 for its workers)."""
    _o().oracle_set_num_threads(int(n))


# --------------------------------------------------------------------------- the real reference
class _Ref:
    def __init__(self):
        self.lib = L = ctypes.CDLL(REF_SO)
        sz = ctypes.c_size_t
        L.ref_brute_force.argtypes = [ctypes.c_int, ctypes.c_int, sz, ctypes.c_void_p, _dp]
        L.ref_brute_force.restype = ctypes.c_double
        L.ref_simulate.argtypes = [ctypes.c_int, sz, ctypes.c_void_p, ctypes.c_double,
                                   ctypes.c_int, ctypes.c_int]
        L.ref_simulate.restype = ctypes.c_int
        L.ref_accuracy_pct.argtypes = [ctypes.c_int, sz, _dp, _dp]
        L.ref_accuracy_pct.restype = ctypes.c_double
        L.ref_bvh_single_leaf_forces.argtypes = [ctypes.c_int, sz, ctypes.c_void_p, _dp]
        L.ref_bvh_single_leaf_forces.restype = ctypes.c_int
        L.ref_bvh_forces.argtypes = [ctypes.c_int, sz, ctypes.c_void_p, _dp]
        L.ref_bvh_forces.restype = ctypes.c_double
        L.ref_G.restype = ctypes.c_double
        L.ref_omp_threads.restype = ctypes.c_int
        L.ref_parlay_workers.restype = ctypes.c_int


_ref = None


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _r():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libnbref.so not built (needs /root/reference)")
        _ref = _Ref()
    return _ref.lib


def ref_forces(bodies, variant="omp_2", want_forces=True):
    """Run the reference's OWN compiled brute_force_<variant>_n_body<D>.  G and the cut-off are
    the reference's compile-time constants.  Returns (forces or None, seconds in the call)."""
    dim = _dim_of(bodies)
    b = _as_bodies(bodies, dim)
    out = np.zeros((b.shape[0], dim)) if want_forces else None
    secs = _r().ref_brute_force(dim, VARIANTS[variant], b.shape[0], b.ctypes.data,
                                _p(out) if want_forces else None)
    if secs < 0:
        raise RuntimeError(f"reference brute force failed ({secs})")
    return out, secs


def ref_simulate(bodies, dt, nsteps, variant="omp_2"):
    dim = _dim_of(bodies)
    b = _as_bodies(bodies, dim).copy()
    rc = _r().ref_simulate(dim, b.shape[0], b.ctypes.data, dt, nsteps, VARIANTS[variant])
    if rc:
        raise RuntimeError(f"reference simulate rc={rc}")
    return b


def ref_bvh_single_leaf_forces(bodies):
    """The reference's own compiled BVH<D>::calculate_force on a tree whose root is one leaf holding every body
    (max_bodies_per_leaf >= n): its leaf loop (bvh.cpp:149-177) as a direct sum over all bodies."""
    dim = _dim_of(bodies)
    b = _as_bodies(bodies, dim)
    out = np.zeros((b.shape[0], dim))
    rc = _r().ref_bvh_single_leaf_forces(dim, b.shape[0], b.ctypes.data, _p(out))
    if rc:
        raise RuntimeError(f"reference BVH leaf forces rc={rc}")
    return out


def ref_bvh_forces(bodies, want_forces=True):
    """The reference's bvh_seq_n_body<D> end to end (leaf size 16): (forces or None, seconds)."""
    dim = _dim_of(bodies)
    b = _as_bodies(bodies, dim)
    out = np.zeros((b.shape[0], dim)) if want_forces else None
    secs = _r().ref_bvh_forces(dim, b.shape[0], b.ctypes.data, _p(out) if want_forces else None)
    if secs < 0:
        raise RuntimeError(f"reference BVH failed ({secs})")
    return out, secs


def ref_accuracy_pct(forces_, reference):
    f = np.ascontiguousarray(forces_, dtype=np.float64)
    r = np.ascontiguousarray(reference, dtype=np.float64)
    return _r().ref_accuracy_pct(f.shape[1], f.shape[0], _p(f), _p(r))


def ref_threads():
    return {"omp": _r().ref_omp_threads(), "parlay": _r().ref_parlay_workers(), "G": _r().ref_G()}
