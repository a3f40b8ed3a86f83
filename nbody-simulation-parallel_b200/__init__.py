"""nbody-simulation-parallel_b200 -- B200-native brute-force N-body path (Python host mirror).

The product is ``lib/libnb200.so`` (hand-written sm_100a kernels behind the C ABI of
``include/nb200.h``) and the C++ adapter ``host/methods_cuda.h`` that gives it the reference's
``methods.h`` signatures.  This module is the thin ctypes mirror used by tests and ``bench.py``;
function names follow the reference's entry points (methods.h:29-43, :85-91):

    brute_force_cuda_n_body(bodies)              -> forces        (cf. brute_force_omp_n_body_2<D>)
    brute_force_cuda_simulate(bodies, dt, steps) -> bodies after  (force + update_body_* loop)

``bodies`` is the reference's AoS ``Body<D>`` as a float64 array (n, 2*D+1).  The directory name
is not a Python identifier; load it with ``__graft_entry__.load_package()``.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from . import generators
from ._lib import NB200_FP32, NB200_FP64

#: the reference's compile-time constants (utils.h:21; methods.cpp:24)
G_REF = 4.471e-21
CUTOFF_REF = 1e-10


#: The FP32-mode parity criterion (include/nb200.h), in ONE place for tests, smoke() and bench.py:
#: per-body norm-wise relative force error against the FP64 reference on the same 24-bit-quantised inputs
#:     err_i <= max(FP32_TOL, FP32_PER_KAPPA * kappa_i),   kappa_i = sum_j |f_ij| / |sum_j f_ij|
#: i.e. the flat 1e-5 of BASELINE.json for every body whose own force sum is not ill-conditioned (kappa <= 16.7) and
#: 10 * 2^-24 * kappa beyond.  The constant comes from full-population measurements (tools/pop_error.py,
#: profiles/r02/pop_error.jsonl): the worst body seen sits at 4.9e-7 * kappa (2D, N=65536, uniform).
FP32_TOL = 1e-5
FP32_PER_KAPPA = 6e-7


def fp32_error_bound(kappa):
    """Per-body FP32-mode error bound for summation condition numbers ``kappa`` (array or scalar)."""
    return np.maximum(FP32_TOL, FP32_PER_KAPPA * np.asarray(kappa, dtype=np.float64))


class NB200Error(RuntimeError):
    pass


def _dim_of(bodies: np.ndarray) -> int:
    if bodies.ndim != 2 or bodies.shape[1] not in (5, 7):
        raise ValueError("bodies must be (n, 5) for 2D or (n, 7) for 3D (Body<D>: pos, vel, mass)")
    return (bodies.shape[1] - 1) // 2


class NBodyCuda:
    """One libnb200 context (``nb200_ctx``).

    ``ngpus`` > 1 drives several GPUs from this process; ``rank``/``world``/``unique_id`` build
    the one-process-per-GPU flavour used under torchrun (see ``distributed.py``).
    """

    def __init__(self, dim: int, n: int, precision: int = NB200_FP64, ngpus: int = 1, *, device: int | None = None,
                 rank: int | None = None, world: int | None = None, unique_id: bytes | None = None):
        self._lib = _lib.load()
        self.dim, self.n, self.precision = dim, n, precision
        self._h = ctypes.c_void_p()
        if rank is None:
            rc = self._lib.nb200_create(ctypes.byref(self._h), dim, n, precision, ngpus)
        else:
            buf = ctypes.create_string_buffer(unique_id, _lib.UNIQUE_ID_BYTES) if unique_id else None
            rc = self._lib.nb200_create_rank(ctypes.byref(self._h), dim, n, precision,
                                             0 if device is None else device, rank, world or 1, buf)
        if rc != 0:
            msg = self._lib.nb200_last_error(None)
            self._h = ctypes.c_void_p()
            raise NB200Error(f"nb200 create failed ({rc}): {msg.decode() if msg else ''}")

    # -- plumbing
    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self._lib.nb200_last_error(self._h)
            raise NB200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.nb200_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ipc_export(self) -> bytes:
        """This rank's CUDA IPC blob for the fused NVLink exchange (nb200_ipc_export)."""
        buf = ctypes.create_string_buffer(_lib.IPC_BYTES)
        self._check(self._lib.nb200_ipc_export(self._h, buf), "ipc_export")
        return buf.raw

    def ipc_attach(self, blobs: list[bytes]):
        """Attach the blobs of ALL ranks, in rank order (nb200_ipc_attach)."""
        joined = b"".join(blobs)
        buf = ctypes.create_string_buffer(joined, len(joined))
        self._check(self._lib.nb200_ipc_attach(self._h, buf, len(blobs)), "ipc_attach")

    def set_option(self, key: str, value: int):
        self._check(self._lib.nb200_set_option(self._h, key.encode(), int(value)), f"set_option({key})")

    # -- data
    def upload(self, bodies: np.ndarray):
        b = np.ascontiguousarray(bodies, dtype=np.float64)
        if b.shape != (self.n, 2 * self.dim + 1):
            raise ValueError(f"expected bodies of shape {(self.n, 2 * self.dim + 1)}, got {b.shape}")
        self._check(self._lib.nb200_upload_aos(self._h, b.ctypes.data, b.strides[0] if self.n else 8 * (2 * self.dim + 1)),
                    "upload")
        self._stride = b.strides[0] if self.n else 8 * (2 * self.dim + 1)

    def generate(self, kind: int, seed: int, G: float = G_REF):
        """Seeded bodies generated on the device (nb200_generate): 0 reference range, 1 uniform, 2 Plummer.
        ``generators.device_bodies`` is the numpy mirror of the same generator."""
        self._check(self._lib.nb200_generate(self._h, int(kind), int(seed), float(G)), "generate")
        self._stride = 8 * (2 * self.dim + 1)

    def download(self, into: np.ndarray) -> np.ndarray:
        """Overwrite the owned rows of ``into`` with the advanced bodies (mass = the uploaded mass)."""
        if into.dtype != np.float64 or not into.flags.c_contiguous or into.shape != (self.n, 2 * self.dim + 1):
            raise ValueError("download target must be a C-contiguous float64 (n, 2D+1) array")
        self._check(self._lib.nb200_download_aos(self._h, into.ctypes.data, self._stride), "download")
        return into

    def shard_range(self) -> tuple[int, int]:
        lo, hi = ctypes.c_size_t(), ctypes.c_size_t()
        self._check(self._lib.nb200_shard_range(self._h, ctypes.byref(lo), ctypes.byref(hi)), "shard_range")
        return lo.value, hi.value

    # -- the hot path
    def forces(self, G: float = G_REF, cutoff: float = CUTOFF_REF, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.zeros((self.n, self.dim))
        self._check(self._lib.nb200_forces(self._h, G, cutoff, out.ctypes.data_as(_lib._dp)), "forces")
        return out

    def step(self, dt: float, nsteps: int = 1, G: float = G_REF, cutoff: float = CUTOFF_REF):
        self._check(self._lib.nb200_step(self._h, G, cutoff, dt, nsteps), "step")

    def energy(self, G: float = G_REF, cutoff: float = CUTOFF_REF) -> tuple[float, float]:
        ke, pe = ctypes.c_double(), ctypes.c_double()
        self._check(self._lib.nb200_energy(self._h, G, cutoff, ctypes.byref(ke), ctypes.byref(pe)), "energy")
        return ke.value, pe.value

    def accuracy_pct(self, forces: np.ndarray | None, reference: np.ndarray) -> float:
        """compute_accuracy_omp (utils.h:170-219) on the device; ``forces=None`` compares the forces of
        the last ``forces()`` call that are still on the device."""
        r = np.ascontiguousarray(reference, dtype=np.float64)
        fp = None
        if forces is not None:
            f = np.ascontiguousarray(forces, dtype=np.float64)
            fp = f.ctypes.data_as(_lib._dp)
        pct = ctypes.c_double()
        self._check(self._lib.nb200_accuracy_pct(self._h, fp, r.ctypes.data_as(_lib._dp), ctypes.byref(pct)),
                    "accuracy_pct")
        return pct.value

    def compare_forces(self, other: "NBodyCuda") -> dict:
        """Per-body norm-wise relative difference between this context's resident forces and ``other``'s
        (the reference side), over ALL bodies, on the device (nb200_compare_forces)."""
        st = (ctypes.c_double * _lib.COMPARE_STATS)()
        self._check(self._lib.nb200_compare_forces(self._h, other._h, st), "compare_forces")
        hist = [int(v) for v in st[4:]]
        edges = [f"<1e-16"] + [f"1e{k - 17}..1e{k - 16}" for k in range(1, 18)]
        return {"bodies": int(st[0]), "max": float(st[1]), "argmax": int(st[2]), "nonfinite": int(st[3]),
                "histogram": {e: c for e, c in zip(edges, hist) if c},
                "over_1e-5": int(sum(hist[12:])), "over_1e-4": int(sum(hist[13:])), "over_1e-3": int(sum(hist[14:]))}

    def validation_forces(self, cap: int = 8) -> tuple[np.ndarray, np.ndarray]:
        """The (index, force) rows print_validation_forces (utils.h:138-151) would print."""
        out = np.zeros((cap, self.dim))
        idx = np.zeros(cap, dtype=np.int64)
        k = self._lib.nb200_validation_forces(self._h, out.ctypes.data_as(_lib._dp),
                                              idx.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), cap)
        if k < 0:
            self._check(k, "validation_forces")
        return idx[:k], out[:k]

    # -- measurement
    @property
    def last_elapsed_ms(self) -> float:
        ms = ctypes.c_double()
        self._check(self._lib.nb200_last_elapsed_ms(self._h, ctypes.byref(ms)), "last_elapsed_ms")
        return ms.value

    @property
    def launch_count(self) -> int:
        return int(self._lib.nb200_launch_count(self._h))

    @property
    def plan(self) -> str:
        p = self._lib.nb200_plan(self._h)
        return p.decode() if p else ""


def measure_fp32_peak(device: int = 0) -> float:
    """Measured FP32 FMA-pipe peak of ``device`` in TFLOP/s (nb200_measure_fp32_peak)."""
    lib = _lib.load()
    t = ctypes.c_double()
    rc = lib.nb200_measure_fp32_peak(device, ctypes.byref(t))
    if rc != 0:
        msg = lib.nb200_last_error(None)
        raise NB200Error(f"nb200_measure_fp32_peak failed ({rc}): {msg.decode() if msg else ''}")
    return t.value


#: the guards of the reference's tree codes for their leaf (P2P) sums
P2P_BVH = {"eps_same": 1e-9, "cutoff": 1e-9, "skip_same_index": 0}       # bvh.cpp:156-169
P2P_FMM = {"eps_same": -1.0, "cutoff": 1e-10, "skip_same_index": 1}      # fmm.cpp:624-628


def p2p_leaves_cuda(bodies: np.ndarray, leaf_offsets, leaf_bodies, nbr_offsets, nbr_leaves, G: float = G_REF,
                    cutoff: float = 1e-9, eps_same: float = 1e-9, skip_same_index: int = 0, sign: int = 1,
                    device: int = 0, return_ms: bool = False):
    """The leaf (P2P) step of the suite's tree codes on the device (nb200_p2p_leaves): every body of target leaf l
    receives the direct sum over the bodies of the leaves ``nbr_leaves[nbr_offsets[l]:nbr_offsets[l+1]]``, with the
    tree codes' attractive sign and guards (defaults: BVH<D>::calculate_force, bvh.cpp:149-177)."""
    b = np.ascontiguousarray(bodies, dtype=np.float64)
    dim = _dim_of(b)
    lo = np.ascontiguousarray(leaf_offsets, dtype=np.int64)
    lb = np.ascontiguousarray(leaf_bodies, dtype=np.int64)
    no = np.ascontiguousarray(nbr_offsets, dtype=np.int64)
    nl = np.ascontiguousarray(nbr_leaves, dtype=np.int64)
    out = np.zeros((b.shape[0], dim))
    ms = ctypes.c_double()
    ll = ctypes.POINTER(ctypes.c_longlong)
    lib = _lib.load()
    rc = lib.nb200_p2p_leaves(device, dim, b.shape[0], b.ctypes.data, b.strides[0] if b.shape[0] else 8 * (2 * dim + 1),
                              max(lo.shape[0] - 1, 0), lo.ctypes.data_as(ll), lb.ctypes.data_as(ll), no.ctypes.data_as(ll),
                              nl.ctypes.data_as(ll), G, cutoff, eps_same, int(skip_same_index), int(sign),
                              out.ctypes.data_as(_lib._dp), ctypes.byref(ms))
    if rc != 0:
        msg = lib.nb200_last_error(None)
        raise NB200Error(f"nb200_p2p_leaves failed ({rc}): {msg.decode() if msg else ''}")
    return (out, ms.value) if return_ms else out


def brute_force_cuda_n_body(bodies: np.ndarray, precision: int = NB200_FP64, ngpus: int = 1, G: float = G_REF,
                            cutoff: float = CUTOFF_REF, options: dict | None = None) -> np.ndarray:
    """Forces on every body, (n, D) float64 in body order -- the ``BruteForce_CUDA`` method beside
    ``brute_force_{seq,omp_1,omp_2,parlay_1,parlay_2}_n_body`` (methods.h:29-43).  Host buffers in,
    host buffers out (the timing convention of ``safely_execute``, utils.h:87-104)."""
    b = np.ascontiguousarray(bodies, dtype=np.float64)
    dim = _dim_of(b)
    with NBodyCuda(dim, b.shape[0], precision, ngpus) as ctx:
        for k, v in (options or {}).items():
            ctx.set_option(k, v)
        ctx.upload(b)
        return ctx.forces(G, cutoff)


def brute_force_cuda_simulate(bodies: np.ndarray, dt: float, steps: int, precision: int = NB200_FP64,
                              ngpus: int = 1, G: float = G_REF, cutoff: float = CUTOFF_REF,
                              options: dict | None = None) -> np.ndarray:
    """``steps`` iterations of {brute force; update_body_velocities; update_body_positions}
    (methods.cpp:426-450) on the device(s); returns the advanced AoS bodies."""
    b = np.array(bodies, dtype=np.float64, order="C", copy=True)
    dim = _dim_of(b)
    with NBodyCuda(dim, b.shape[0], precision, ngpus) as ctx:
        for k, v in (options or {}).items():
            ctx.set_option(k, v)
        ctx.upload(b)
        ctx.step(dt, steps, G, cutoff)
        ctx.download(b)
    return b


__all__ = ["NBodyCuda", "NB200Error", "NB200_FP32", "NB200_FP64", "G_REF", "CUTOFF_REF", "FP32_TOL", "FP32_PER_KAPPA",
           "fp32_error_bound",
           "measure_fp32_peak", "brute_force_cuda_n_body", "brute_force_cuda_simulate", "p2p_leaves_cuda", "P2P_BVH", "P2P_FMM",
           "generators"]
