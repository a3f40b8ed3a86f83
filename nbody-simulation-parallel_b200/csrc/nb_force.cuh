// nb_force.cuh -- the brute-force all-pairs force pass with the integrator fused into its epilogue.
//
// Replaces (reference, /root/reference/nbody-sim-new): the pair loop of brute_force_*_n_body<D>
// (methods.cpp:7-42, :98-136: d = p_j - p_i; skip if r^2 < 1e-10; F_i -= G m_i m_j d / r^4) and the
// two integrator helpers update_body_velocities (:426-438, v += (F/m) dt) and
// update_body_positions (:441-450, x += v dt).
//
// Work decomposition (persistent CTAs, dynamic units):
//   unit = (i-tile of ITILE own targets) x (segment of seg_tiles source tiles)
//   CTA  = BLOCK threads; JS adjacent lanes share TI register-resident targets and split the
//          sources of a tile between them; partial sums are reduced with warp shuffles, added
//          into FP64 global accumulators, and the CTA that finishes the LAST unit of an i-tile
//          runs the epilogue for it (forces out, or v/x update + next-step source tile).
//   Source tiles stream through a NB_STAGES-deep shared-memory ring filled by 1-D TMA bulk
//   copies (cp.async.bulk + mbarrier), one copy per 256-source tile.
//
// FP32 pair arithmetic is written in packed f32x2 (FADD2/FMUL2/FFMA2, sm_100+): two SOURCES per
// instruction against one duplicated target, so the per-pair FMA-pipe work (11 lane-ops in 3D)
// costs 5.5 issue slots and MUFU.RCP + the cut-off select fit in the free slots.  No sqrt/rsqrt
// is needed: the reference law is d / r^4 = d * (1/r^2)^2.  Tensor cores are not used: this is
// not a contraction.
#pragma once
#include "nb_common.cuh"

template <bool F64> struct NbReal { using type = float; };
template <> struct NbReal<true> { using type = double; };

// ---------------------------------------------------------------------------------------------
// FP32: one tile against TI targets.  npos holds the NEGATED target coordinates (the packed add
// takes them as a broadcast scalar operand).  a[][] receives this tile's FP32 partial sums.
//
// The loop is register-file bound on sm_100 (one operand fetch per clock per SM sub-partition:
// 25 fetches per packed chain without any cut-off work, 22 FMA-pipe clocks), so every extra
// instruction in it costs its operand count.  Three flavours:
// NB_EXACT   hard cut-off per pair (methods.cpp:119): pairs with r^2 < cutoff are DROPPED, which
//            also removes the self pair and exact duplicates: rcp(+inf) = 0.  FSETP+FSEL per pair.
// NB_TRACKED no per-pair cut-off work; tracks the minimum r^2 seen (one FMNMX3 per two pairs) and
//            the caller REDOES the tile with NB_EXACT when that minimum is under the cut-off (the
//            tile holding the thread's own targets, duplicates, genuinely close pairs).  NaN/inf
//            produced by rcp(0) in such a pass are discarded with the rest of it.
// NB_PLAIN   nothing at all; legal only where a pre-pass (nb_grid_*_kernel) proved that no target
//            of the warp has another body within the cut-off radius and the tile does not hold
//            the warp's own targets.
enum { NB_PLAIN = 0, NB_TRACKED = 1, NB_EXACT = 2 };

#define NB_STR_(x) #x
#define NB_UNROLL(n) _Pragma(NB_STR_(unroll n))

// EQM (equal-mass system, see nb_force_sym.cuh): the sums are taken WITHOUT the source masses (s = 1/r^4), the common
// mass is applied once per body in the epilogue.
// HL (48-bit positions, see nb_force_sym.cuh): lo planes behind the D + 1 hi planes of the stage, nlo = negated lo parts
// of the targets.
template <int D, int TI, int JS, int MODE, int UNR, bool EQM = false, bool HL = false>
__device__ __forceinline__ float nb_tile_f32(const float* __restrict__ stage, int part, float cutoff,
                                             const float (&npos)[TI][3], float2 (&a)[TI][3],
                                             const float (*nlo)[3] = nullptr) {
    const float4* sx = reinterpret_cast<const float4*>(stage);
    const float4* sy = sx + NB_TILE / 4;
    const float4* sz = sy + NB_TILE / 4;                      // D == 3 only
    const float4* sm = sx + D * (NB_TILE / 4);
    const float4* lx = sm + NB_TILE / 4;                      // HL only
    const float4* ly = lx + NB_TILE / 4;
    const float4* lz = ly + NB_TILE / 4;
#pragma unroll
    for (int t = 0; t < TI; ++t)
#pragma unroll
        for (int d = 0; d < 3; ++d) a[t][d] = make_float2(0.f, 0.f);

    const float inf = __int_as_float(0x7f800000);
    float rmin = inf;
    NB_UNROLL(UNR)
    for (int q = part; q < NB_TILE / 4; q += JS) {
        const float4 X = sx[q], Y = sy[q], M = sm[q];
        float4 Z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (D == 3) Z = sz[q];
        float4 XL = make_float4(0.f, 0.f, 0.f, 0.f), YL = XL, ZL = XL;
        if (HL) {
            XL = lx[q];
            YL = ly[q];
            if (D == 3) ZL = lz[q];
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float2 xs = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
            const float2 ys = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
            const float2 zs = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
            const float2 ms = h ? make_float2(M.z, M.w) : make_float2(M.x, M.y);
            const float2 xl = h ? make_float2(XL.z, XL.w) : make_float2(XL.x, XL.y);
            const float2 yl = h ? make_float2(YL.z, YL.w) : make_float2(YL.x, YL.y);
            const float2 zl = h ? make_float2(ZL.z, ZL.w) : make_float2(ZL.x, ZL.y);
#pragma unroll
            for (int t = 0; t < TI; ++t) {
                float2 dx = __fadd2_rn(xs, make_float2(npos[t][0], npos[t][0]));
                float2 dy = __fadd2_rn(ys, make_float2(npos[t][1], npos[t][1]));
                if (HL) {
                    dx = __fadd2_rn(dx, __fadd2_rn(xl, make_float2(nlo[t][0], nlo[t][0])));
                    dy = __fadd2_rn(dy, __fadd2_rn(yl, make_float2(nlo[t][1], nlo[t][1])));
                }
                float2 r2 = __fmul2_rn(dx, dx);
                r2 = __ffma2_rn(dy, dy, r2);
                float2 dz;
                if (D == 3) {
                    dz = __fadd2_rn(zs, make_float2(npos[t][2], npos[t][2]));
                    if (HL) dz = __fadd2_rn(dz, __fadd2_rn(zl, make_float2(nlo[t][2], nlo[t][2])));
                    r2 = __ffma2_rn(dz, dz, r2);
                }
                if (MODE == NB_EXACT) {
                    r2.x = (r2.x >= cutoff) ? r2.x : inf;
                    r2.y = (r2.y >= cutoff) ? r2.y : inf;
                } else if (MODE == NB_TRACKED) {
                    rmin = fminf(rmin, fminf(r2.x, r2.y));    // one FMNMX3
                }
                float2 inv;
                inv.x = nb_rcp_f32(r2.x);
                inv.y = nb_rcp_f32(r2.y);
                float2 s = __fmul2_rn(inv, inv);
                if (!EQM) s = __fmul2_rn(s, ms);
                a[t][0] = __ffma2_rn(s, dx, a[t][0]);
                a[t][1] = __ffma2_rn(s, dy, a[t][1]);
                if (D == 3) a[t][2] = __ffma2_rn(s, dz, a[t][2]);
            }
        }
    }
    return rmin;
}

// FP64: one tile against TI targets (pos holds the plain target coordinates).
template <int D, int TI, int JS, bool EXACT, bool EQM = false>
__device__ __forceinline__ void nb_tile_f64(const double* __restrict__ stage, int part, double cutoff,
                                            const double (&pos)[TI][3], double (&accd)[TI][3]) {
    const double2* sx = reinterpret_cast<const double2*>(stage);
    const double2* sy = sx + NB_TILE / 2;
    const double2* sz = sy + NB_TILE / 2;                     // D == 3 only
    const double2* sm = sx + D * (NB_TILE / 2);
#pragma unroll 2
    for (int q = part; q < NB_TILE / 2; q += JS) {
        const double2 X = sx[q], Y = sy[q], M = sm[q];
        double2 Z = make_double2(0.0, 0.0);
        if (D == 3) Z = sz[q];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double xs = h ? X.y : X.x, ys = h ? Y.y : Y.x, zs = h ? Z.y : Z.x;
            const double ms = h ? M.y : M.x;
#pragma unroll
            for (int t = 0; t < TI; ++t) {
                const double dx = xs - pos[t][0];
                const double dy = ys - pos[t][1];
                double r2 = dx * dx;
                r2 = fma(dy, dy, r2);
                double dz = 0.0;
                if (D == 3) {
                    dz = zs - pos[t][2];
                    r2 = fma(dz, dz, r2);
                }
                double inv = nb_rcp_f64(r2);
                if (EXACT) inv = (r2 >= cutoff) ? inv : 0.0;  // drop (also kills the NaN of r2 = 0)
                const double s = EQM ? inv * inv : ms * (inv * inv);
                accd[t][0] = fma(s, dx, accd[t][0]);
                accd[t][1] = fma(s, dy, accd[t][1]);
                if (D == 3) accd[t][2] = fma(s, dz, accd[t][2]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// FLAGS = true: P.suspect[] (from the close-pair pre-pass) selects NB_PLAIN / NB_EXACT per warp and
// tile; FLAGS = false: self-contained NB_TRACKED pass with redo (FP32) or NB_EXACT always (FP64).
template <int D, bool F64, int TI, int JS, int BLOCK, bool FLAGS, int UNR = 2>
__global__ void __launch_bounds__(BLOCK) nb_force_kernel(const NbForceParams P) {
    using real = typename NbReal<F64>::type;
    constexpr int NP = D + 1;                       // planes per tile
    constexpr int GROUPS = BLOCK / JS;              // target groups per CTA
    constexpr int ITILE = GROUPS * TI;              // targets per i-tile
    constexpr int TILE_ELEMS = NB_TILE * NP;
    constexpr uint32_t TILE_BYTES = TILE_ELEMS * sizeof(real);
    constexpr int NWARPS = BLOCK / 32;
    static_assert(JS >= 1 && JS <= 32 && (JS & (JS - 1)) == 0, "JS must be a power of two <= 32");

    extern __shared__ __align__(128) unsigned char nb_smem[];
    real* ring = reinterpret_cast<real*>(nb_smem);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(nb_smem + (size_t)NB_STAGES * TILE_BYTES);
    uint64_t* empty_bar = full_bar + NB_STAGES;
    int* s_unit = reinterpret_cast<int*>(empty_bar + NB_STAGES);   // [0] unit, [1] last-flag

    const int tid = threadIdx.x;
    const int group = tid / JS;
    const int part = tid % JS;
    const real* __restrict__ src = static_cast<const real*>(P.src);

    if (tid == 0) {
        for (int s = 0; s < NB_STAGES; ++s) {
            nb_mbar_init(&full_bar[s], 1);
            nb_mbar_init(&empty_bar[s], NWARPS);
        }
        nb_fence_mbar_init();
    }
    __syncthreads();

    const float cutoff_f = (float)P.cutoff;
    // redo threshold of the FP32 fast pass: the cut-off, but never 0 (r^2 = 0 must always redo)
    const float cutoff_redo = fmaxf(cutoff_f, 1.0e-37f);
    // cross-GPU handshake: this pass reads rows that the peers' epilogues of the previous step wrote
    // into OUR source buffer over NVLink; wait until every peer has published that step.
    bool need_wait = (P.wait_step | P.wait_epoch) != 0ull && P.n_peers > 0;
    auto peer_handshake = [&]() {
        if (tid < P.n_peers) {
            const unsigned long long* f = P.my_flags + P.peer_rank[tid];
            nb_wait_flag(f, P.wait_step, P.spin_timeout_ns, P.err_word, NB_WAIT_STEP, P.peer_rank[tid]) &&
                nb_wait_flag(f + P.flag_stride, P.wait_epoch, P.spin_timeout_ns, P.err_word, NB_WAIT_EPOCH, P.peer_rank[tid]);
        }
        __syncthreads();
    };
    if (need_wait && !P.lazy_wait) {
        peer_handshake();
        need_wait = false;
    }

    const int total_units = P.n_itiles * (P.rng_nseg[0] + P.rng_nseg[1] + P.rng_nseg[2]);
    unsigned kt = 0;                                 // tiles consumed by this CTA so far (ring position)

    for (;;) {
        if (tid == 0) s_unit[0] = (int)atomicAdd(&P.sched[0], 1u);
        __syncthreads();
        const int u = s_unit[0];
        __syncthreads();
        if (u >= total_units) break;
        const int it = u % P.n_itiles;
        const int seg = u / P.n_itiles;
        int rsel = 0, sloc = seg;
        if (sloc >= P.rng_nseg[0]) { sloc -= P.rng_nseg[0]; rsel = 1; }
        if (rsel == 1 && sloc >= P.rng_nseg[1]) { sloc -= P.rng_nseg[1]; rsel = 2; }
        const int ts = P.rng_begin[rsel] + sloc * P.seg_tiles;
        const int te = min(ts + P.seg_tiles, P.rng_end[rsel]);
        const int ntl = te - ts;
        if (need_wait && rsel != 0) {          // first remote unit of this CTA (u is CTA-uniform)
            peer_handshake();
            need_wait = false;
        }

        // producer prologue: up to NB_STAGES-1 tiles in flight before the first wait
        if (tid == 0) {
            const int pre = min(NB_STAGES - 1, ntl);
            for (int t = 0; t < pre; ++t) {
                const unsigned k = kt + t;
                const int slot = k % NB_STAGES;
                nb_mbar_wait(&empty_bar[slot], ((k / NB_STAGES) & 1u) ^ 1u);
                nb_mbar_expect_tx(&full_bar[slot], TILE_BYTES);
                nb_tma_load_1d(ring + (size_t)slot * TILE_ELEMS, src + (size_t)(ts + t) * TILE_ELEMS,
                               TILE_BYTES, &full_bar[slot]);
            }
        }

        // register-resident targets of this thread's group
        real tpos[TI][3];
        int own_tile[TI];
        bool suspect = false;
#pragma unroll
        for (int t = 0; t < TI; ++t) {
            const long long b = P.tgt_base + (long long)it * ITILE + group + t * GROUPS;
            own_tile[t] = (int)(b / NB_TILE);
            const real* tb = src + (size_t)(b / NB_TILE) * TILE_ELEMS + (b % NB_TILE);
#pragma unroll
            for (int d = 0; d < 3; ++d) tpos[t][d] = (d < D) ? tb[d * NB_TILE] : real(0);
            if constexpr (FLAGS) suspect |= P.suspect[it * ITILE + group + t * GROUPS] != 0;
        }
        const bool warp_suspect = FLAGS ? (__any_sync(0xffffffffu, suspect) != 0) : false;
        double accd[TI][3];
#pragma unroll
        for (int t = 0; t < TI; ++t)
#pragma unroll
            for (int d = 0; d < 3; ++d) accd[t][d] = 0.0;

        float npos[TI][3];
        if (!F64) {
#pragma unroll
            for (int t = 0; t < TI; ++t)
#pragma unroll
                for (int d = 0; d < 3; ++d) npos[t][d] = -(float)tpos[t][d];
        }

        for (int t = 0; t < ntl; ++t) {
            if (tid == 0 && t + NB_STAGES - 1 < ntl) {
                const unsigned k = kt + t + NB_STAGES - 1;
                const int slot = k % NB_STAGES;
                nb_mbar_wait(&empty_bar[slot], ((k / NB_STAGES) & 1u) ^ 1u);
                nb_mbar_expect_tx(&full_bar[slot], TILE_BYTES);
                nb_tma_load_1d(ring + (size_t)slot * TILE_ELEMS,
                               src + (size_t)(ts + t + NB_STAGES - 1) * TILE_ELEMS, TILE_BYTES,
                               &full_bar[slot]);
            }
            const unsigned k = kt + t;
            const int slot = k % NB_STAGES;
            nb_mbar_wait(&full_bar[slot], (k / NB_STAGES) & 1u);
            const real* stage = ring + (size_t)slot * TILE_ELEMS;
            // with FLAGS: exact pass for warps with a suspect target and for the tiles that hold the
            // warp's own targets (self pairs); both conditions are warp-uniform
            bool exact_tile = true;
            if constexpr (FLAGS) {
                exact_tile = warp_suspect;
#pragma unroll
                for (int tt = 0; tt < TI; ++tt) exact_tile |= (ts + t == own_tile[tt]);
            }
            if constexpr (F64) {
                const double* dstage = reinterpret_cast<const double*>(stage);
                if (exact_tile)
                    nb_tile_f64<D, TI, JS, true>(dstage, part, P.cutoff, reinterpret_cast<const double(&)[TI][3]>(tpos), accd);
                else
                    nb_tile_f64<D, TI, JS, false>(dstage, part, P.cutoff, reinterpret_cast<const double(&)[TI][3]>(tpos), accd);
            } else {
                const float* fstage = reinterpret_cast<const float*>(stage);
                float2 a[TI][3];
                if constexpr (FLAGS) {
                    if (exact_tile) nb_tile_f32<D, TI, JS, NB_EXACT, UNR>(fstage, part, cutoff_f, npos, a);
                    else nb_tile_f32<D, TI, JS, NB_PLAIN, UNR>(fstage, part, cutoff_f, npos, a);
                } else {
                    const float rmin = nb_tile_f32<D, TI, JS, NB_TRACKED, UNR>(fstage, part, cutoff_f, npos, a);
                    if (!(rmin >= cutoff_redo)) nb_tile_f32<D, TI, JS, NB_EXACT, UNR>(fstage, part, cutoff_f, npos, a);
                }
                // per-tile flush of the short FP32 partial sums into FP64
#pragma unroll
                for (int tt = 0; tt < TI; ++tt)
#pragma unroll
                    for (int d = 0; d < D; ++d) accd[tt][d] += (double)(a[tt][d].x + a[tt][d].y);
            }
            __syncwarp();
            if ((tid & 31) == 0) nb_mbar_arrive(&empty_bar[slot]);
        }
        kt += ntl;

        // reduce the JS source-parts of each target with warp shuffles, then one FP64 atomic per
        // (target, component) into the global accumulators
#pragma unroll
        for (int t = 0; t < TI; ++t)
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double v = accd[t][d];
#pragma unroll
                for (int o = JS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (part == 0) {
                    const int li = it * ITILE + group + t * GROUPS;
                    if (P.slots) __stcg(&P.slots[((size_t)seg * 3 + d) * P.tpad + li], v);
                    else atomicAdd(&P.acc[(size_t)d * P.tpad + li], v);
                }
            }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned old = atomicAdd(&P.tile_done[it], 1u);
            s_unit[1] = (old + 1u == P.units_per_itile);
        }
        __syncthreads();
        if (s_unit[1]) {
            // ---------------- fused epilogue: this CTA finished the last unit of i-tile `it`
            __threadfence();
            for (int k = tid; k < ITILE; k += BLOCK) {
                const int li = it * ITILE + k;
                double S[3];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    if (P.slots) {
                        double v = 0.0;                    // fixed order: segment 0, 1, 2, ...
                        for (int sg = 0; sg < P.nseg_total; ++sg) v += __ldcg(&P.slots[((size_t)sg * 3 + d) * P.tpad + li]);
                        S[d] = v * P.acc_scale;
                    } else {
                        double* ap = &P.acc[(size_t)d * P.tpad + li];
                        S[d] = __ldcg(ap) * P.acc_scale;
                        __stcg(ap, 0.0);                   // self-clean for the next step
                    }
                }
                if (li >= P.n_local) continue;
                const double m = P.mass[li];
                const double gm = P.G * m;
                double F[3];
#pragma unroll
                for (int d = 0; d < D; ++d) F[d] = -(gm * S[d]);     // forces[i] -= ... (methods.cpp:131)
                if (P.mode == 0) {
#pragma unroll
                    for (int d = 0; d < D; ++d) P.forces[(size_t)li * D + d] = F[d];
                } else {
                    const long long b = P.tgt_base + li;
                    real* nb = static_cast<real*>(P.src_next) + (size_t)(b / NB_TILE) * TILE_ELEMS + (b % NB_TILE);
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        double v = P.vel[(size_t)d * P.tpad + li];
                        double x = P.pos[(size_t)d * P.tpad + li];
                        v += (F[d] / m) * P.dt;               // methods.cpp:436
                        x += v * P.dt;                        // methods.cpp:448 (uses the NEW v)
                        P.vel[(size_t)d * P.tpad + li] = v;
                        P.pos[(size_t)d * P.tpad + li] = x;
                        const real xs = (real)(x * P.pos_scale);
                        nb[d * NB_TILE] = xs;
                        // fused all-gather: the same row goes to every peer's next buffer (P2P store)
                        const size_t off = (size_t)(b / NB_TILE) * TILE_ELEMS + (b % NB_TILE) + (size_t)d * NB_TILE;
                        for (int pr = 0; pr < P.n_peers; ++pr) static_cast<real*>(P.peer_next[pr])[off] = xs;
                    }
                }
            }
            if (P.n_peers > 0 && P.mode == 1) __threadfence_system();   // remote rows before the exit count
            if (tid == 0) P.tile_done[it] = 0u;
        }
    }

    // last CTA out resets the dynamic scheduler for the next launch
    if (tid == 0) {
        const unsigned e = atomicAdd(&P.sched[1], 1u);
        if (e + 1u == gridDim.x) {
            P.sched[0] = 0u;
            P.sched[1] = 0u;
            __threadfence();
            if (P.signal_step != 0ull) {
                // every CTA fenced its remote rows (system scope) before its exit count above
                __threadfence_system();
                for (int pr = 0; pr < P.n_peers; ++pr) nb_st_release_sys(P.peer_flags[pr] + P.my_rank, P.signal_step);
            }
        }
    }
}
