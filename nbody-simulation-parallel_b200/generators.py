"""Seeded synthetic body sets (harness inputs; numpy only, no GPU, no oracle).

The reference's only generator is unseeded (``generate_random_bodies<D>``, utils.h:107-135:
``std::random_device``), so identical bytes cannot be fed to two implementations with it.  These
generators reproduce its RANGES where the config asks for them and add the two distributions
BASELINE.json names (uniform cube, Plummer sphere).  Output layout is the reference's AoS
``Body<D>`` (body.h:7-19): float64 array (n, 2*D+1) = position[D], velocity[D], mass.
"""
from __future__ import annotations

import numpy as np

#: utils.h:21
G_REF = 4.471e-21


def reference_range(n: int, dim: int = 3, seed: int = 42) -> np.ndarray:
    """The reference generator's ranges (utils.h:113-115), seeded: pos U[1,1e7], vel U[-10,10],
    mass U[1,1e8]."""
    rng = np.random.default_rng(seed)
    b = np.empty((n, 2 * dim + 1))
    b[:, :dim] = rng.uniform(1.0, 1.0e7, (n, dim))
    b[:, dim:2 * dim] = rng.uniform(-10.0, 10.0, (n, dim))
    b[:, 2 * dim] = rng.uniform(1.0, 1.0e8, n)
    return b


def uniform_cube(n: int, dim: int = 3, seed: int = 42, G: float = G_REF) -> np.ndarray:
    """Unit cube/square: pos U[0,1)^D, vel U[-0.1,0.1], mass U[0.5,1.5]/(G n) so that
    G * sum(m) ~ 1 and the dynamics are non-trivial with the reference's tiny G."""
    rng = np.random.default_rng(seed)
    b = np.empty((n, 2 * dim + 1))
    b[:, :dim] = rng.random((n, dim))
    b[:, dim:2 * dim] = rng.uniform(-0.1, 0.1, (n, dim))
    b[:, 2 * dim] = rng.uniform(0.5, 1.5, n) / (G * max(n, 1))
    return b


def jittered_cube(n: int, dim: int = 3, seed: int = 42, G: float = G_REF, fill: float = 0.5) -> np.ndarray:
    """Stratified uniform cube/square: one body per cell of a k^D lattice (k = ceil(n^(1/D)), a random
    subset of n cells), placed uniformly inside the central ``fill`` fraction of its cell, so that no
    two bodies are closer than (1-fill)/k.  Same density, velocities and masses as ``uniform_cube``.
    Poisson-uniform points in 2D hold pairs at r ~ 1/n whose r^-3 forces no fixed time step
    resolves; energy-drift comparisons use this set instead."""
    rng = np.random.default_rng(seed)
    k = int(np.ceil(n ** (1.0 / dim) - 1e-9))
    cells = rng.permutation(k ** dim)[:n]
    b = np.empty((n, 2 * dim + 1))
    for d in range(dim):
        c = (cells // (k ** d)) % k
        b[:, d] = (c + 0.5 + fill * (rng.random(n) - 0.5)) / k
    b[:, dim:2 * dim] = rng.uniform(-0.1, 0.1, (n, dim))
    b[:, 2 * dim] = rng.uniform(0.5, 1.5, n) / (G * max(n, 1))
    return b


def plummer(n: int, seed: int = 42, G: float = G_REF, a: float = 1.0, rmax: float = 22.8) -> np.ndarray:
    """3D Plummer sphere (scale a), equal masses 1/(G n): r = a (u^(-2/3) - 1)^(-1/2) truncated
    at rmax*a, isotropic directions; speeds by the standard q^2 (1-q^2)^(7/2) rejection times the
    local escape speed sqrt(2) (1 + r^2)^(-1/4)."""
    rng = np.random.default_rng(seed)
    r = np.empty(n)
    filled = 0
    while filled < n:
        u = rng.random(n - filled)
        u = u[u > 0]
        rr = a / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
        rr = rr[rr < rmax * a]
        r[filled:filled + rr.size] = rr
        filled += rr.size

    def iso(k):
        z = rng.uniform(-1.0, 1.0, k)
        phi = rng.uniform(0.0, 2.0 * np.pi, k)
        s = np.sqrt(1.0 - z * z)
        return np.stack([s * np.cos(phi), s * np.sin(phi), z], axis=1)

    q = np.empty(n)
    filled = 0
    while filled < n:
        k = n - filled
        x = rng.random(2 * k)
        y = rng.random(2 * k) * 0.1
        ok = x[y < x * x * (1.0 - x * x) ** 3.5][:k]
        q[filled:filled + ok.size] = ok
        filled += ok.size
    speed = q * np.sqrt(2.0) * (1.0 + (r / a) ** 2) ** (-0.25)
    b = np.empty((n, 7))
    b[:, 0:3] = iso(n) * r[:, None]
    b[:, 3:6] = iso(n) * speed[:, None]
    b[:, 6] = 1.0 / (G * max(n, 1))
    return b


# ---------------------------------------------------------------------------------------------
# numpy mirror of the DEVICE generator (csrc/nb_aux.cuh: NbPhilox, nb_generate_kernel), so that the
# oracle can be fed the very bodies nb200_generate produced without a download.
def _philox_uniforms(seed: int, index: np.ndarray, stream: int | np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Two 53-bit uniforms in [0,1) per (index, stream): Philox-4x32-10, key = seed, counter = (index, stream, 0)."""
    u32 = np.uint64(0xFFFFFFFF)
    idx = np.asarray(index, dtype=np.uint64)
    c0 = idx & u32
    c1 = idx >> np.uint64(32)
    c2 = np.broadcast_to(np.asarray(stream, dtype=np.uint64), idx.shape).copy()
    c3 = np.zeros_like(idx)
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    for _ in range(10):
        p0, p1 = m0 * c0, m1 * c2                      # 32x32 -> 64-bit products, exact in uint64
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & u32, p1 >> np.uint64(32), p1 & u32
        c0, c1, c2, c3 = h1 ^ c1 ^ k0, l1, h0 ^ c3 ^ k1, l0
        k0 = (k0 + np.uint64(0x9E3779B9)) & u32
        k1 = (k1 + np.uint64(0xBB67AE85)) & u32
    two53 = 2.0 ** -53
    a = (((c0 << np.uint64(32)) | c1) >> np.uint64(11)).astype(np.float64) * two53
    b = (((c2 << np.uint64(32)) | c3) >> np.uint64(11)).astype(np.float64) * two53
    return a, b


def device_bodies(n: int, dim: int, kind: int, seed: int, G: float = G_REF) -> np.ndarray:
    """The bodies ``NBodyCuda.generate(kind, seed, G)`` creates on the device.  Kinds 0 (reference
    range) and 1 (uniform) match bit for bit; kind 2 (Plummer) to the last ulps of pow/sin/cos."""
    idx = np.arange(n, dtype=np.uint64)
    u = np.empty((8, n))
    for k in range(4):
        u[2 * k], u[2 * k + 1] = _philox_uniforms(seed, idx, k)
    b = np.empty((n, 2 * dim + 1))

    def lerp(lo, hi, x):
        return lo + (hi - lo) * x

    if kind == 0:
        for d in range(dim):
            b[:, d] = lerp(1.0, 1.0e7, u[d])
            b[:, dim + d] = lerp(-10.0, 10.0, u[3 + d])
        b[:, 2 * dim] = lerp(1.0, 1.0e8, u[6])
    elif kind == 1:
        for d in range(dim):
            b[:, d] = u[d]
            b[:, dim + d] = lerp(-0.1, 0.1, u[3 + d])
        b[:, 2 * dim] = lerp(0.5, 1.5, u[6]) / (G * float(n))
    elif kind == 2:
        if dim != 3:
            raise ValueError("the Plummer generator is 3D only")
        r = np.zeros(n)
        todo = np.arange(n)
        att = 0
        while todo.size:
            a0, _ = _philox_uniforms(seed, todo.astype(np.uint64), 16 + att)
            with np.errstate(divide="ignore", invalid="ignore"):
                rr = 1.0 / np.sqrt(a0 ** (-2.0 / 3.0) - 1.0)
            ok = (a0 > 0.0) & (rr < 22.8)
            r[todo[ok]] = rr[ok]
            todo = todo[~ok]
            att += 1
        q = np.zeros(n)
        todo = np.arange(n)
        att = 0
        while todo.size:
            a0, a1 = _philox_uniforms(seed, todo.astype(np.uint64), 1024 + att)
            ok = 0.1 * a1 < a0 * a0 * (1.0 - a0 * a0) ** 3.5
            q[todo[ok]] = a0[ok]
            todo = todo[~ok]
            att += 1
        speed = q * np.sqrt(2.0) * (1.0 + r * r) ** -0.25
        two_pi = 6.283185307179586
        for col, (uz, up, mag) in enumerate(((u[0], u[1], r), (u[2], u[3], speed))):
            z = 2.0 * uz - 1.0
            ph = two_pi * up
            sq = np.sqrt(1.0 - z * z)
            b[:, 3 * col + 0] = sq * np.cos(ph) * mag
            b[:, 3 * col + 1] = sq * np.sin(ph) * mag
            b[:, 3 * col + 2] = z * mag
        b[:, 6] = 1.0 / (G * float(n))
    else:
        raise ValueError("kind must be 0 (reference range), 1 (uniform) or 2 (Plummer)")
    return b


def round_to_float(bodies: np.ndarray) -> np.ndarray:
    """Positions and masses rounded to float32 and widened back: what the FP32 pair kernel sees.
    The <=1e-5 FP32 criterion compares against the FP64 oracle fed THESE inputs."""
    dim = (bodies.shape[1] - 1) // 2
    b = np.array(bodies, dtype=np.float64, copy=True)
    b[:, :dim] = b[:, :dim].astype(np.float32).astype(np.float64)
    b[:, 2 * dim] = b[:, 2 * dim].astype(np.float32).astype(np.float64)
    return b


def relative_norm_error(f: np.ndarray, ref: np.ndarray) -> np.ndarray:
    """Per-body norm-wise relative error ||f_i - ref_i||_2 / ||ref_i||_2 (SURVEY section 4 lesson i).
    Bodies whose reference force is exactly zero compare absolutely (error 0 iff f_i == 0)."""
    num = np.linalg.norm(np.asarray(f) - np.asarray(ref), axis=1)
    den = np.linalg.norm(np.asarray(ref), axis=1)
    out = np.where(den > 0, num / np.where(den > 0, den, 1.0), np.where(num > 0, np.inf, 0.0))
    return out
