// nb_p2p.cuh -- direct (particle-to-particle) sums over LEAF LISTS: the P2P step of the suite's tree codes.
//
// Replaces (reference, /root/reference/nbody-sim-new): the leaf branch of BVH<D>::calculate_force (bvh.cpp:149-177),
// the leaf branch of FMM<D>::calculate_accurate_force (fmm.cpp:622-637) and the direct-evaluation loops of
// fmm_parlay.cpp:918-1023 -- the SAME pair law as the brute-force methods (|F| = G m_i m_j / r^3 along d = p_j - p_i,
// i.e. F = G m_i m_j d / r^4) with the tree codes' ATTRACTIVE sign (force += diff.normalized() * mag) and their
// guards: BVH skips a pair when every |d_k| <= 1e-9 and when r^2 < 1e-9; FMM skips the body itself (pointer equality)
// and pairs with r^2 < 1e-10.  Sign, guards and G are run-time parameters.
//
// Layout: the bodies are gathered into LEAF ORDER on the device (planar FP64: x | y | (z) | m, leaf after leaf), so a
// leaf is a contiguous run.  One CTA per target leaf: lane = target (up to 32 at a time; smaller leaves give the spare
// lanes a share of the sources), the CTA's warps split the source leaves of the leaf's neighbour list between them and
// stage each through shared memory; per-target partial sums of the warps meet in shared memory and the force goes to its body's row of the output (a body is the target of
// exactly one leaf: no atomics).  All FP64: tree-code P2P is held to the same 1e-12 as the brute-force path.
#pragma once
#include "nb_common.cuh"

#define NB_P2P_BLOCK 128
#define NB_P2P_CHUNK 64                   // sources staged per warp at a time

struct NbP2PParams {
    const double* lpos;                   // [D][total] leaf-ordered coordinates
    const double* lmass;                  // [total]
    const long long* lbody;               // [total] body index of every leaf-ordered slot
    const long long* leaf_off;            // [n_leaves + 1]
    const long long* nbr_off;             // [n_leaves + 1]
    const long long* nbr_leaf;            // source leaf ids
    double* forces;                       // [n][D] body order
    long long total;
    long long n_leaves;
    double G, cutoff, eps_same;           // eps_same < 0: no per-component "same position" test
    int skip_same_index;                  // 1: never pair a slot with the same BODY (FMM's other == &body)
    double sign;                          // +1 attractive (tree codes), -1 the brute-force methods' convention
};

// AoS bodies -> leaf-ordered planar arrays
template <int D>
__global__ void __launch_bounds__(256) nb_p2p_gather_kernel(const double* __restrict__ aos, size_t stride_d,
                                                             const long long* __restrict__ lbody, long long total,
                                                             double* __restrict__ lpos, double* __restrict__ lmass) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const double* rec = aos + (size_t)lbody[k] * stride_d;
#pragma unroll
    for (int d = 0; d < D; ++d) lpos[(size_t)d * total + k] = rec[d];
    lmass[k] = rec[2 * D];
}

template <int D>
__global__ void __launch_bounds__(NB_P2P_BLOCK) nb_p2p_leaf_kernel(const NbP2PParams P) {
    constexpr int NW = NB_P2P_BLOCK / 32;
    __shared__ double s_src[NW][D + 1][NB_P2P_CHUNK];
    __shared__ long long s_body[NW][NB_P2P_CHUNK];
    __shared__ double s_sum[NW][D][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long leaf = blockIdx.x; leaf < P.n_leaves; leaf += gridDim.x) {
        const long long t0 = P.leaf_off[leaf], t1 = P.leaf_off[leaf + 1];
        const long long nb0 = P.nbr_off[leaf], nb1 = P.nbr_off[leaf + 1];
        // a group of `tw` lanes (a power of two >= the targets at hand, at most 32) holds one target per lane; the 32 / tw
        // groups of the warp split the staged sources between them (a 16-body leaf keeps both half-warps busy)
        for (long long tb = t0; tb < t1; tb += 32) {
            const int nt = (int)min(32LL, t1 - tb);
            int tw = 1;
            while (tw < nt) tw <<= 1;
            const int js = 32 / tw, part = lane / tw;
            const long long ti = tb + (lane & (tw - 1));
            const bool live = ti < t1;
            double xi[D], mi = 0.0, acc[D];
            long long bi = -1;
#pragma unroll
            for (int d = 0; d < D; ++d) { xi[d] = live ? P.lpos[(size_t)d * P.total + ti] : 0.0; acc[d] = 0.0; }
            if (live) { mi = P.lmass[ti]; bi = P.lbody[ti]; }
            // the warps take the source leaves of the neighbour list round-robin
            for (long long q = nb0 + warp; q < nb1; q += NW) {
                const long long sl = P.nbr_leaf[q];
                const long long s0 = P.leaf_off[sl], s1 = P.leaf_off[sl + 1];
                for (long long sb = s0; sb < s1; sb += NB_P2P_CHUNK) {
                    const int cnt = (int)min((long long)NB_P2P_CHUNK, s1 - sb);
                    __syncwarp();
                    for (int k = lane; k < cnt; k += 32) {
#pragma unroll
                        for (int d = 0; d < D; ++d) s_src[warp][d][k] = P.lpos[(size_t)d * P.total + sb + k];
                        s_src[warp][D][k] = P.lmass[sb + k];
                        s_body[warp][k] = P.lbody[sb + k];
                    }
                    __syncwarp();
                    for (int k = part; k < cnt; k += js) {
                        double dd[D], r2 = 0.0;
                        bool same = P.eps_same >= 0.0;
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            dd[d] = s_src[warp][d][k] - xi[d];
                            r2 = fma(dd[d], dd[d], r2);
                            same = same && !(fabs(dd[d]) > P.eps_same);           // bvh.cpp:156-163
                        }
                        const bool self = P.skip_same_index && s_body[warp][k] == bi;   // fmm.cpp:624
                        const bool keep = !same && !self && r2 >= P.cutoff;          // bvh.cpp:169 / fmm.cpp:628
                        const double inv = nb_rcp_f64(keep ? r2 : 1.0);
                        const double w = keep ? inv * inv * s_src[warp][D][k] : 0.0;
#pragma unroll
                        for (int d = 0; d < D; ++d) acc[d] = fma(w, dd[d], acc[d]);
                    }
                }
            }
            // groups -> the group of part 0 (every lane takes part: the loop bounds above are lane dependent)
            for (int o = tw; o < 32; o <<= 1) {
#pragma unroll
                for (int d = 0; d < D; ++d) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
            }
            // warps -> one sum per target
#pragma unroll
            for (int d = 0; d < D; ++d) s_sum[warp][d][lane] = acc[d];
            __syncthreads();
            if (warp == 0 && live && part == 0) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    double v = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) v += s_sum[w][d][lane];
                    P.forces[(size_t)bi * D + d] = P.sign * (P.G * mi) * v;
                }
            }
            __syncthreads();
        }
    }
}
