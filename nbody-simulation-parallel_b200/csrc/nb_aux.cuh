// nb_aux.cuh -- layout conversion and diagnostics kernels around the force pass.
#pragma once
#include "nb_common.cuh"

// Equal-mass systems run the pair kernels without the masses (nb_force_sym.cuh, EQM), so a padding body cannot be
// silenced by its zero mass: it is parked out of range instead, where 1/r^4 underflows to exactly 0 against every real
// body (FP32: r^2 ~ 1e30, FP64: 1e300; distinct per body, so padding never coincides with padding).
template <typename real> __device__ __forceinline__ real nb_park_coordinate(long long b) {
    const double far = sizeof(real) == 8 ? 1.0e150 : 1.0e15;
    return (real)(far * (1.5 + (double)(b & 1023) / 1024.0));     // never the 1.0 * far the force kernel parks idle lanes at
}

// AoS Body<D> (body.h:7-19; stride_d doubles per body) -> tile-planar sources for ALL bodies of
// both ring buffers, plus the FP64 master state of the own targets.  Padding bodies (index >= n)
// get mass 0 and the position of body 0: every pair with them contributes exactly 0.
template <int D, typename real>
__global__ void nb_pack_kernel(const double* __restrict__ aos, size_t stride_d, long long n,
                               long long nalloc, real* __restrict__ src0, real* __restrict__ src1,
                               double pos_scale, double mass_scale, long long tgt_base, int tpad,
                               double* __restrict__ pos, double* __restrict__ vel,
                               double* __restrict__ mass, int park) {
    constexpr int NP = D + 1;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nalloc) return;
    const bool real_body = b < n;
    const double* rec = aos + (size_t)(real_body ? b : 0) * stride_d;
    real* d0 = src0 + (size_t)(b / NB_TILE) * (NB_TILE * NP) + (b % NB_TILE);
    real* d1 = src1 + (size_t)(b / NB_TILE) * (NB_TILE * NP) + (b % NB_TILE);
    double x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        x[d] = (n > 0) ? rec[d] : 0.0;
        const real xs = (park && !real_body) ? nb_park_coordinate<real>(b) : (real)(x[d] * pos_scale);
        d0[d * NB_TILE] = xs;
        d1[d * NB_TILE] = xs;
    }
    const double m = real_body ? rec[2 * D] : 0.0;
    const real ms = (real)(m * mass_scale);
    d0[D * NB_TILE] = ms;
    d1[D * NB_TILE] = ms;
    const long long li = b - tgt_base;
    if (li >= 0 && li < tpad) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            pos[(size_t)d * tpad + li] = x[d];
            vel[(size_t)d * tpad + li] = real_body ? rec[D + d] : 0.0;
        }
        mass[li] = m;
    }
}

// 48-bit positions (option "fp32_positions" = 48): the part of every scaled coordinate the float row loses, as a second
// float, tile-planar [D][256] per tile, both buffers.  Padding bodies get 0.
template <int D>
__global__ void nb_pack_lo_kernel(const double* __restrict__ aos, size_t stride_d, long long n, long long nalloc,
                                  float* __restrict__ lo0, float* __restrict__ lo1, double pos_scale) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nalloc) return;
    const size_t off = (size_t)(b / NB_TILE) * (NB_TILE * D) + (b % NB_TILE);
#pragma unroll
    for (int d = 0; d < D; ++d) {
        float lo = 0.f;
        if (b < n) {
            const double xs = aos[(size_t)b * stride_d + d] * pos_scale;
            lo = (float)(xs - (double)(float)xs);
        }
        lo0[off + (size_t)d * NB_TILE] = lo;
        lo1[off + (size_t)d * NB_TILE] = lo;
    }
}

// Shard-local flavour (multi-GPU with the peer-store exchange): the image holds only the rows this shard owns; their
// tile-planar source rows go to BOTH buffers of this shard and of every peer (plain stores on peer-mapped pointers over
// NVLink), so no rank ever uploads or packs a body it does not own.  Thread t < tpad: own padded row t (rows past the
// shard's tiles only reset the master state); t >= tpad: one of the slack rows behind the last tile (local only).
struct NbPeerBufs {
    void* buf0[NB_MAX_PEERS];
    void* buf1[NB_MAX_PEERS];
    int n_peers;
};
template <int D, typename real>
__global__ void nb_pack_shard_kernel(const double* __restrict__ aos, size_t stride_d, long long n, long long tgt_base,
                                     int span, int tpad, long long nbodies, long long nalloc,
                                     real* __restrict__ src0, real* __restrict__ src1, NbPeerBufs peers,
                                     double pos_scale, double mass_scale, double* __restrict__ pos,
                                     double* __restrict__ vel, double* __restrict__ mass, int park) {
    constexpr int NP = D + 1;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool slack = t >= tpad;
    const long long b = slack ? nbodies + (t - tpad) : tgt_base + t;
    if (slack && b >= nalloc) return;
    const bool own_row = !slack && t < span;
    const bool real_body = own_row && b < n;
    // padding rows sit on the shard's first body (or the origin of an empty shard) with mass 0: they add exactly 0
    const bool have_anchor = tgt_base < n;
    const double* rec = aos + (size_t)(real_body ? b : (have_anchor ? tgt_base : 0)) * stride_d;
    double x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = (real_body || have_anchor) ? rec[d] : 0.0;
    const double m = real_body ? rec[2 * D] : 0.0;
    if (own_row || slack) {
        const size_t off = (size_t)(b / NB_TILE) * (NB_TILE * NP) + (b % NB_TILE);
#pragma unroll
        for (int d = 0; d <= D; ++d) {
            const real v = d < D ? ((park && !real_body) ? nb_park_coordinate<real>(b) : (real)(x[d] * pos_scale))
                                 : (real)(m * mass_scale);
            src0[off + (size_t)d * NB_TILE] = v;
            src1[off + (size_t)d * NB_TILE] = v;
            if (own_row)
                for (int p = 0; p < peers.n_peers; ++p) {
                    static_cast<real*>(peers.buf0[p])[off + (size_t)d * NB_TILE] = v;
                    static_cast<real*>(peers.buf1[p])[off + (size_t)d * NB_TILE] = v;
                }
        }
    }
    if (!slack) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            pos[(size_t)d * tpad + t] = own_row ? x[d] : 0.0;
            vel[(size_t)d * tpad + t] = real_body ? rec[D + d] : 0.0;
        }
        mass[t] = m;
    }
}

// max |coordinate|, max |mass| and min |mass| over the AoS image (FP32 mode picks its power-of-two source
// scales from the first two; min == max marks an equal-mass system).  Non-negative doubles order like their bit patterns, so the reduction ends
// in one 64-bit atomicMax per warp.  NaNs are ignored (every comparison with them is false).
template <int D>
__global__ void __launch_bounds__(256) nb_bounds_kernel(const double* __restrict__ aos, size_t stride_d,
                                                         long long n, unsigned long long* __restrict__ out) {
    double xm = 0.0, mm = 0.0, mlo = __longlong_as_double(0x7ff0000000000000ll);      // +inf
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < n;
         b += (long long)gridDim.x * blockDim.x) {
        const double* rec = aos + (size_t)b * stride_d;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double v = fabs(rec[d]);
            if (v > xm) xm = v;
        }
        const double m = fabs(rec[2 * D]);
        if (m > mm) mm = m;
        if (m < mlo) mlo = m;
    }
    unsigned long long xb = (unsigned long long)__double_as_longlong(xm);
    unsigned long long mb = (unsigned long long)__double_as_longlong(mm);
    unsigned long long lb = (unsigned long long)__double_as_longlong(mlo);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long x2 = __shfl_xor_sync(0xffffffffu, xb, o);
        const unsigned long long m2 = __shfl_xor_sync(0xffffffffu, mb, o);
        const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lb, o);
        xb = x2 > xb ? x2 : xb;
        mb = m2 > mb ? m2 : mb;
        lb = l2 < lb ? l2 : lb;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&out[0], xb);
        atomicMax(&out[1], mb);
        atomicMin(&out[2], lb);              // out[2] starts at all ones
    }
}

// master state of the own targets -> rows [tgt_base, tgt_base + n_local) of an AoS device image
template <int D>
__global__ void nb_unpack_kernel(double* __restrict__ aos_rows, size_t stride_d, long long n_local,
                                 int tpad, const double* __restrict__ pos,
                                 const double* __restrict__ vel) {
    const long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_local) return;
    double* rec = aos_rows + (size_t)li * stride_d;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        rec[d] = pos[(size_t)d * tpad + li];
        rec[D + d] = vel[(size_t)d * tpad + li];
    }
}

// Energy of the reference's own law in FP64: per own target i
//   ke_i = 1/2 m_i v_i^2,   pe_i = (G m_i / 4) * sum_{j != i, r^2 >= cutoff} m_j / r^2
// (pe summed over ordered pairs, hence 1/4 = 1/2 * 1/2).  Sources come from the tile-planar buffer (float-rounded
// in FP32 mode, undone by inv_pos_scale / inv_mass_scale).  Same scaffold as the force pass on a small scale: a
// source tile is staged once per CTA in its storage type, every thread keeps TWO targets in registers, the
// reciprocal is MUFU.RCP64H + one cubic step (nb_rcp_f64) instead of a true divide, and the pair loop is 8 DP
// operations per pair (3 DADD/DFMA for r^2 in 3D, 3 for 1/r^2, 1 DFMA for the sum, 1 DSETP for the cut-off).
#define NB_ENERGY_TI 2
template <int D, typename real>
__global__ void __launch_bounds__(256) nb_energy_kernel(const real* __restrict__ src, long long ntiles,
                                                         long long tgt_base, long long n_local, int tpad,
                                                         const double* __restrict__ vel,
                                                         const double* __restrict__ mass, double G,
                                                         double cutoff_scaled, double inv_pos_scale,
                                                         double inv_mass_scale,
                                                         double* __restrict__ out /* [2] ke, pe */) {
    constexpr int NP = D + 1;
    __shared__ real tile[NB_TILE * NP];
    __shared__ double red[2][8];
    double xi[NB_ENERGY_TI][D], sum[NB_ENERGY_TI];
    long long li[NB_ENERGY_TI];
#pragma unroll
    for (int t = 0; t < NB_ENERGY_TI; ++t) {
        li[t] = ((long long)blockIdx.x * NB_ENERGY_TI + t) * blockDim.x + threadIdx.x;
        const long long b = tgt_base + (li[t] < n_local ? li[t] : 0);
        const real* tb = src + (size_t)(b / NB_TILE) * (NB_TILE * NP) + (b % NB_TILE);
#pragma unroll
        for (int d = 0; d < D; ++d) xi[t][d] = (double)tb[d * NB_TILE];
        sum[t] = 0.0;
    }
    for (long long tl = 0; tl < ntiles; ++tl) {
        __syncthreads();
        for (int k = threadIdx.x; k < NB_TILE * NP; k += blockDim.x) tile[k] = src[(size_t)tl * (NB_TILE * NP) + k];
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < NB_TILE; ++j) {
            double xj[D];
#pragma unroll
            for (int d = 0; d < D; ++d) xj[d] = (double)tile[d * NB_TILE + j];
            const double mj = (double)tile[D * NB_TILE + j];
#pragma unroll
            for (int t = 0; t < NB_ENERGY_TI; ++t) {
                double r2 = 0.0;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const double dd = xj[d] - xi[t][d];
                    r2 = fma(dd, dd, r2);
                }
                const bool keep = r2 >= cutoff_scaled;              // drops the self pair and duplicates as well (r2 = 0)
                const double inv = nb_rcp_f64(keep ? r2 : 1.0);
                sum[t] = fma(keep ? mj : 0.0, inv, sum[t]);
            }
        }
    }
    double ke = 0.0, pe = 0.0;
#pragma unroll
    for (int t = 0; t < NB_ENERGY_TI; ++t) {
        if (li[t] < n_local) {
            const double m = mass[li[t]];
            double v2 = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const double v = vel[(size_t)d * tpad + li[t]];
                v2 = fma(v, v, v2);
            }
            ke += 0.5 * m * v2;
            // sum is in scaled units: m' / r'^2 = (m ms) / (r^2 ps^2)
            pe += 0.25 * G * m * sum[t] * inv_mass_scale / (inv_pos_scale * inv_pos_scale);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ke += __shfl_xor_sync(0xffffffffu, ke, o);
        pe += __shfl_xor_sync(0xffffffffu, pe, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = ke;
        red[1][threadIdx.x >> 5] = pe;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            a += red[0][w];
            c += red[1][w];
        }
        atomicAdd(&out[0], a);
        atomicAdd(&out[1], c);
    }
}

// ---------------------------------------------------------------------------------------------
// Close-pair pre-pass (O(N)): lets the force kernel drop ALL per-pair cut-off work (NB_PLAIN).
// A uniform hash grid with cell edge h >= sqrt(cutoff) * 1.001 counts the bodies per cell; a
// target is "suspect" when the 3^D cells around it hold any body besides itself.  Every pair
// with r^2 < cutoff (as the kernel computes it: FP32 rounding moves r^2 by < 1e-6 relative)
// lies in adjacent cells, so both of its bodies are flagged; false positives only cost speed.
struct NbGrid {
    unsigned long long* keys;   // open addressing, linear probing; ~0ull = empty
    unsigned* counts;           // bodies in the cell MINUS ONE: one memset of 0xFF clears keys and counts together
    unsigned mask;              // capacity - 1 (power of two)
    double inv_h;               // 1 / cell edge, in source units
};

__device__ __forceinline__ unsigned long long nb_cell_key(long long cx, long long cy, long long cz) {
    unsigned long long k = (unsigned long long)cx * 0x9E3779B97F4A7C15ull;
    k ^= ((unsigned long long)cy * 0xC2B2AE3D27D4EB4Full) + 0x165667B19E3779F9ull + (k << 6) + (k >> 2);
    k ^= ((unsigned long long)cz * 0xD6E8FEB86659FD93ull) + 0x9E3779B97F4A7C15ull + (k << 6) + (k >> 2);
    return k == ~0ull ? 0ull : k;
}
__device__ __forceinline__ unsigned nb_slot_of(unsigned long long key, unsigned mask) {
    return (unsigned)((key * 0xFF51AFD7ED558CCDull) >> 32) & mask;
}

template <int D, typename real>
__global__ void nb_grid_insert_kernel(const real* __restrict__ src, long long nbodies, NbGrid g) {
    // nbodies: every source incl. the zero-mass padding that sits on a real body's position; with parked padding
    // (equal-mass systems) only the real bodies
    constexpr int NP = D + 1;
    nb_launch_dependents();     // the query kernel may queue up behind this one (it waits for its completion itself)
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbodies) return;
    const real* tb = src + (size_t)(b / NB_TILE) * (NB_TILE * NP) + (b % NB_TILE);
    long long c[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < D; ++d) c[d] = (long long)floor((double)tb[d * NB_TILE] * g.inv_h);
    const unsigned long long key = nb_cell_key(c[0], c[1], c[2]);
    unsigned slot = nb_slot_of(key, g.mask);
    for (;;) {
        const unsigned long long prev = atomicCAS(&g.keys[slot], ~0ull, key);
        if (prev == ~0ull || prev == key) break;
        slot = (slot + 1) & g.mask;
    }
    atomicAdd(&g.counts[slot], 1u);
}

template <int D, typename real>
__global__ void nb_grid_query_kernel(const real* __restrict__ src, long long tgt_base, int tpad,
                                     long long nbodies, NbGrid g, unsigned char* __restrict__ suspect) {
    constexpr int NP = D + 1;
    nb_grid_dep_wait();         // every insert is complete and visible
    nb_launch_dependents();     // the force pass may queue up
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= tpad) return;
    const long long b = tgt_base + li;
    if (b >= nbodies) { suspect[li] = 1; return; }      // slack rows past the last source tile
    const real* tb = src + (size_t)(b / NB_TILE) * (NB_TILE * NP) + (b % NB_TILE);
    long long c[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < D; ++d) c[d] = (long long)floor((double)tb[d * NB_TILE] * g.inv_h);
    unsigned total = 0;
    for (int dz = (D == 3 ? -1 : 0); dz <= (D == 3 ? 1 : 0); ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const unsigned long long key = nb_cell_key(c[0] + dx, c[1] + dy, c[2] + dz);
                unsigned slot = nb_slot_of(key, g.mask);
                for (;;) {
                    const unsigned long long k = g.keys[slot];
                    if (k == key) { total += g.counts[slot] + 1u; break; }
                    if (k == ~0ull) break;
                    slot = (slot + 1) & g.mask;
                }
            }
    suspect[li] = total > 1u;                             // anything besides the body itself
}

// ---------------------------------------------------------------------------------------------
// Seeded body generators on the device.  The reference's generator
// (generate_random_bodies<D>, utils.h:107-135) is unseeded; these reproduce its RANGES
// (kind 0: pos U[1,1e7], vel U[-10,10], mass U[1,1e8], utils.h:113-115) and add the two
// distributions BASELINE.json names (kind 1: uniform cube/square, kind 2: Plummer sphere), keyed by
// (seed, body index) through Philox-4x32-10, so every shard/rank generates identical bytes and the
// host mirror in generators.py (numpy) reproduces them for the oracle.  Output: the AoS Body<D>
// image (position[D], velocity[D], mass) that nb200_upload_aos would have copied in.
struct NbPhilox {
    unsigned k0, k1;
    __host__ __device__ static void mulhilo(unsigned a, unsigned b, unsigned& hi, unsigned& lo) {
        const unsigned long long p = (unsigned long long)a * b;
        hi = (unsigned)(p >> 32);
        lo = (unsigned)p;
    }
    // two 53-bit uniforms in [0,1) from counter (index, stream)
    __host__ __device__ void draw(unsigned long long index, unsigned stream, double& u0, double& u1) const {
        unsigned c0 = (unsigned)index, c1 = (unsigned)(index >> 32), c2 = stream, c3 = 0u;
        unsigned a = k0, b = k1;
        for (int r = 0; r < 10; ++r) {
            unsigned h0, l0, h1, l1;
            mulhilo(0xD2511F53u, c0, h0, l0);
            mulhilo(0xCD9E8D57u, c2, h1, l1);
            const unsigned n0 = h1 ^ c1 ^ a, n1 = l1, n2 = h0 ^ c3 ^ b, n3 = l0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        u0 = (double)((((unsigned long long)c0 << 32) | c1) >> 11) * 0x1.0p-53;
        u1 = (double)((((unsigned long long)c2 << 32) | c3) >> 11) * 0x1.0p-53;
    }
};

__device__ __forceinline__ double nb_lerp(double lo, double hi, double u) {
    return __dadd_rn(lo, __dmul_rn(hi - lo, u));          // no FMA contraction: bit-identical to the host mirror
}

template <int D>
__global__ void __launch_bounds__(256) nb_generate_kernel(double* __restrict__ aos, size_t stride_d, long long n,
                                                           int kind, unsigned long long seed, double G) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    NbPhilox rng{(unsigned)seed, (unsigned)(seed >> 32)};
    double u[8];
    for (int k = 0; k < 4; ++k) rng.draw((unsigned long long)i, (unsigned)k, u[2 * k], u[2 * k + 1]);
    double* rec = aos + (size_t)i * stride_d;
    if (kind == 0) {
        for (int d = 0; d < D; ++d) rec[d] = nb_lerp(1.0, 1.0e7, u[d]);
        for (int d = 0; d < D; ++d) rec[D + d] = nb_lerp(-10.0, 10.0, u[3 + d]);
        rec[2 * D] = nb_lerp(1.0, 1.0e8, u[6]);
    } else if (kind == 1) {
        for (int d = 0; d < D; ++d) rec[d] = u[d];
        for (int d = 0; d < D; ++d) rec[D + d] = nb_lerp(-0.1, 0.1, u[3 + d]);
        rec[2 * D] = __ddiv_rn(nb_lerp(0.5, 1.5, u[6]), __dmul_rn(G, (double)n));
    } else {
        // Plummer sphere, scale 1, truncated at 22.8: r = (u^(-2/3) - 1)^(-1/2); speeds by the standard
        // q^2 (1 - q^2)^(7/2) rejection times the escape speed sqrt(2) (1 + r^2)^(-1/4); equal masses
        double r = 0.0, q = 0.0, a0, a1;
        for (unsigned att = 0;; ++att) {
            rng.draw((unsigned long long)i, 16u + att, a0, a1);
            if (a0 > 0.0) {
                r = 1.0 / sqrt(pow(a0, -2.0 / 3.0) - 1.0);
                if (r < 22.8) break;
            }
        }
        for (unsigned att = 0;; ++att) {
            rng.draw((unsigned long long)i, 1024u + att, a0, a1);
            if (0.1 * a1 < a0 * a0 * pow(1.0 - a0 * a0, 3.5)) { q = a0; break; }
        }
        const double speed = q * sqrt(2.0) * pow(1.0 + r * r, -0.25);
        const double two_pi = 6.283185307179586;
        const double z0 = 2.0 * u[0] - 1.0, p0 = two_pi * u[1], s0 = sqrt(1.0 - z0 * z0);
        const double z1 = 2.0 * u[2] - 1.0, p1 = two_pi * u[3], s1 = sqrt(1.0 - z1 * z1);
        const double dir0[3] = {s0 * cos(p0), s0 * sin(p0), z0}, dir1[3] = {s1 * cos(p1), s1 * sin(p1), z1};
        for (int d = 0; d < D; ++d) {
            rec[d] = dir0[d] * r;
            rec[D + d] = dir1[d] * speed;
        }
        rec[2 * D] = __ddiv_rn(1.0, __dmul_rn(G, (double)n));
    }
}

// ---------------------------------------------------------------------------------------------
// compute_accuracy_omp<D> (utils.h:170-219) on the device: a body counts as accurate when EVERY
// component is within 1 % of the reference component; reference components under 1e-20 in
// magnitude are held to |force| <= 1e-9 instead (utils.h:25-26, :191-197).
template <int D>
__global__ void __launch_bounds__(256) nb_accuracy_kernel(const double* __restrict__ forces,
                                                           const double* __restrict__ reference,
                                                           long long n, unsigned long long* __restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool good = false;
    if (i < n) {
        good = true;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double r = reference[i * D + d], f = forces[i * D + d];
            if (fabs(r) < 1e-20) good = good && !(fabs(f) > 1e-9);
            else good = good && !(fabs((f - r) / r) > 0.01);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, good);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}

// ---------------------------------------------------------------------------------------------
// Per-body norm-wise relative difference ||F_a - F_b|| / ||F_b|| of two force arrays resident on the
// device, over ALL bodies: maximum (value and body), and a histogram by decade.  This is the
// full-population form of the parity metric (tests/test_gpu_parity.py) for sizes where the CPU
// oracle can only cover sampled targets: FP32 mode against an FP64 context holding the same bodies.
//   out[0] = bit pattern of the max (non-negative doubles order like their bits), out[1] = body of the max,
//   out[2 + k] = bodies with difference in [10^(k-17), 10^(k-16)), k = 1..17; k = 0: below 1e-16 (or both zero);
//   out[2 + 18] = non-finite differences
#define NB_CMP_BINS 19
template <int D>
__global__ void __launch_bounds__(256) nb_compare_kernel(const double* __restrict__ fa, const double* __restrict__ fb,
                                                          long long n, long long index_base, int pass,
                                                          unsigned long long* __restrict__ out) {
    __shared__ unsigned hist[NB_CMP_BINS];
    if (threadIdx.x < NB_CMP_BINS) hist[threadIdx.x] = 0u;
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double num = 0.0, den = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double a = fa[i * D + d], b = fb[i * D + d];
            num = fma(a - b, a - b, num);
            den = fma(b, b, den);
        }
        const double rel = (num == 0.0) ? 0.0 : sqrt(num / den);      // den = 0 with num > 0: inf
        const unsigned long long bits = (unsigned long long)__double_as_longlong(rel);
        if (pass == 0) {
            int bin;
            if (!isfinite(rel)) bin = NB_CMP_BINS - 1;
            else if (rel < 1e-16) bin = 0;
            else bin = min(NB_CMP_BINS - 2, max(1, (int)floor(log10(rel)) + 17));
            atomicAdd(&hist[bin], 1u);
            if (isfinite(rel)) atomicMax(&out[0], bits);
        } else if (isfinite(rel) && bits == out[0]) {
            atomicMin(&out[1], (unsigned long long)(index_base + i));     // pass 1: the first body that attains the maximum
        }
    }
    __syncthreads();
    if (pass == 0 && threadIdx.x < NB_CMP_BINS && hist[threadIdx.x])
        atomicAdd(&out[2 + threadIdx.x], (unsigned long long)hist[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// Measured FP32 FMA-pipe peak of the device the roofline is quoted against (MEASURED_PEAKS.json
// holds only HBM and bf16-tensor peaks): independent packed FFMA2 chains acc = x*x + acc, eight per
// thread, two operands each so the register file is not the limit (tools/ubench.cu: 97 % of
// SMs x 128 lanes x 2 x clock).
#define NB_PEAK_ITERS 4096
#define NB_PEAK_NACC 8
__global__ void __launch_bounds__(256) nb_fma_peak_kernel(float* __restrict__ out, float a) {
    float2 acc[NB_PEAK_NACC], x[NB_PEAK_NACC];
#pragma unroll
    for (int i = 0; i < NB_PEAK_NACC; ++i) {
        acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
        x[i] = make_float2(a + i * 1e-6f, a - i * 1e-6f);
    }
#pragma unroll 1
    for (int it = 0; it < NB_PEAK_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NB_PEAK_NACC; ++i) acc[i] = __ffma2_rn(x[i], x[i], acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NB_PEAK_NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
