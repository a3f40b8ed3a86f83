// nb_aux.cuh -- layout conversion and diagnostics kernels around the force pass.
#pragma once
#include "nb_common.cuh"

// AoS Body<D> (body.h:7-19; stride_d doubles per body) -> tile-planar sources for ALL bodies of
// both ring buffers, plus the FP64 master state of the own targets.  Padding bodies (index >= n)
// get mass 0 and the position of body 0: every pair with them contributes exactly 0.
template <int D, typename real>
__global__ void nb_pack_kernel(const double* __restrict__ aos, size_t stride_d, long long n,
                               long long nalloc, real* __restrict__ src0, real* __restrict__ src1,
                               double pos_scale, double mass_scale, long long tgt_base, int tpad,
                               double* __restrict__ pos, double* __restrict__ vel,
                               double* __restrict__ mass) {
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
        const real xs = (real)(x[d] * pos_scale);
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
// (pe summed over ordered pairs, hence 1/4 = 1/2 * 1/2).  Sources come from the tile-planar
// buffer (float-rounded in FP32 mode, undone by inv_pos_scale / inv_mass_scale).
template <int D, typename real>
__global__ void __launch_bounds__(256) nb_energy_kernel(const real* __restrict__ src, long long ntiles,
                                                         long long tgt_base, long long n_local, int tpad,
                                                         const double* __restrict__ vel,
                                                         const double* __restrict__ mass, double G,
                                                         double cutoff_scaled, double inv_pos_scale,
                                                         double inv_mass_scale,
                                                         double* __restrict__ out /* [2] ke, pe */) {
    constexpr int NP = D + 1;
    __shared__ double tile[NB_TILE * NP];
    __shared__ double red[2][8];
    const long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = li < n_local;
    const long long b = tgt_base + (active ? li : 0);
    double xi[D];
    const real* tb = src + (size_t)(b / NB_TILE) * (NB_TILE * NP) + (b % NB_TILE);
#pragma unroll
    for (int d = 0; d < D; ++d) xi[d] = (double)tb[d * NB_TILE];
    double sum = 0.0;
    for (long long t = 0; t < ntiles; ++t) {
        __syncthreads();
        for (int k = threadIdx.x; k < NB_TILE * NP; k += blockDim.x)
            tile[k] = (double)src[(size_t)t * (NB_TILE * NP) + k];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < NB_TILE; ++j) {
            double r2 = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const double dd = tile[d * NB_TILE + j] - xi[d];
                r2 = fma(dd, dd, r2);
            }
            const double inv = (r2 >= cutoff_scaled) ? 1.0 / r2 : 0.0;
            sum = fma(tile[D * NB_TILE + j], inv, sum);
        }
    }
    double ke = 0.0, pe = 0.0;
    if (active) {
        const double m = mass[li];
        double v2 = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double v = vel[(size_t)d * tpad + li];
            v2 = fma(v, v, v2);
        }
        ke = 0.5 * m * v2;
        // sum is in scaled units: m' / r'^2 = (m ms) / (r^2 ps^2)
        pe = 0.25 * G * m * sum * inv_mass_scale / (inv_pos_scale * inv_pos_scale);
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
