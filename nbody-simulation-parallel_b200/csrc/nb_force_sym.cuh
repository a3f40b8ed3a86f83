// nb_force_sym.cuh -- the pair-symmetric flavour of the force pass (FP32 and FP64).
//
// The reference's sequential variant walks only j > i and scatters +-force to both bodies
// (brute_force_seq_n_body, methods.cpp:18-39: forces[j] += f; forces[i] -= f).  This kernel does
// the same on the GPU: every unordered pair is evaluated once and feeds two accumulators, so the
// shared part of the chain (3 FADD2 + FMUL2 + 2 FFMA2 for d and r^2, MUFU.RCP, FMUL2 for 1/r^4) is
// paid once per two interactions: 15 packed FMA-pipe instructions per two pairs of bodies = 7.5
// lane-ops per ordered interaction instead of 11 (FP64: 18 DP operations per pair instead of 28).
//
// Work list (built on the host, nb200_api.cu build_sym_rows): rows = (i-tile of 1024 own
// targets) x (run of source tiles), cut into units of seg_tiles tiles.  Per i-tile one ORDERED row
// (its own four source tiles, evaluated as ordered pairs by the tile functions of nb_force.cuh:
// self pairs, exact cut-off) and one SYMMETRIC row (the own tiles above it); with several ranks
// also the cross-shard blocks this rank is responsible for (its targets x another shard's sources).
//
// A thread keeps TI targets in registers ("a" side, summed per tile into FP64 exactly like
// nb_force_kernel); the reaction on the streamed sources ("b" side) is produced per lane for four
// sources (FP64: two) per iteration and reduced
//   lane partials -> warp  : ALGO 1/2 (default): the sums of a group of sources travel through the warp in registers,
//                            one lane per step (nb_tile_f32_sym_rot / nb_tile_f64_sym_rot, 12 SHFL per step);
//                            ALGO 0 (round 1): a 32 x 12-word shared-memory transpose (6 STS.64, 16 LDS, 16 FADD)
//   warp sums     -> CTA   : per-warp [D][256] tile buffers, summed by the CTA after the tile
//   CTA tile sums -> global: one FP64 atomicAdd per (source, component) per tile, into an
//                            accumulator array indexed by GLOBAL body (any body can receive)
// Flavours: EQM (equal-mass systems: the masses leave the chain), HL (FP32 on 48-bit positions: float pairs hi + lo).
// The integrator/forces epilogue cannot be fused here (an i-tile also receives "b" sums from the
// units of other i-tiles and, across ranks, from other GPUs: nb_sym_push_kernel), so
// nb_finish_kernel runs after the pass.
#pragma once
#include <algorithm>

#include "nb_force.cuh"

#define NB_SYM_ITILE 1024                           // targets per i-tile (TI * BLOCK) of the large shapes = 4 source tiles;
                                                    // the small-N shape (4 targets x 64 threads) uses 256 = one source tile
#define NB_SYM_ROW 14                               // floats per lane row of the transpose scratch (12 used):
                                                    // 14 makes the STS.64 writes and the column reads bank-conflict free

// One row of the work list: an i-tile of own targets against a run of source tiles.
//   NB_ROW_SYM set:   every pair of the block is evaluated once and feeds both bodies
//   NB_ROW_SYM clear: ordered pairs, targets only (the diagonal block holding the i-tile itself)
struct NbSymRow {
    int it;                      // own i-tile (1024 targets)
    int t_begin, t_end;          // source tiles [t_begin, t_end), global tile indices
    int flags;
};
#define NB_ROW_SYM 1

// sub-tiles per source tile a symmetric row's units are measured in
__host__ __device__ constexpr int nb_sym_subtiles(bool f64, int algo) { return algo == 0 ? 1 : f64 ? 4 : 2; }

// Measured and rejected variants of the rotation flavour (profiles/r02/sym_variants_*.jsonl, small_n.md): step loop
// unrolled twice (-4 %), source pairs loaded half a step ahead (-8 %: loop-carried loads turn into MOVs), the accumulate
// FFMA2 ordered for operand reuse in the source (ptxas reschedules them: no effect), __shfl_sync instead of the
// in-place shfl.sync below (-4 %: 14 MOVs per step), FP64 per-target sums in registers instead of shared memory (-4 %).

struct NbSymParams {
    const void* src;             // tile-planar sources (current step), all bodies (float or double)
    const float* src_lo;         // HL flavour: the lo parts of the scaled coordinates, tile-planar [D][256] per tile
    double* gacc;                // [3][gstride] FP64 accumulators indexed by GLOBAL body index
    size_t gstride;
    unsigned* sched;             // [2] unit counter + exit counter (self-resetting)
    const NbSymRow* rows;        // [n_rows]
    const int* row_prefix;       // [n_rows + 1] first flat unit index of every row
    const unsigned char* suspect;   // [tpad] close-pair flags of the own targets; null = no pre-pass ran: exact cut-off everywhere
    long long tgt_base;          // first own body (multiple of NB_TILE)
    int own_count;               // bodies of this shard incl. tile padding; targets past it are inert
    int n_rows;
    int seg_sub;                 // SYMMETRIC rows: sub-tiles per unit (a tile = SUBT sub-tiles: 1 transpose flavour,
                                 // 2 FP32 rotation (128 sources), 4 FP64 rotation (64 sources))
    int seg_ord;                 // ORDERED rows: source tiles per unit
    int total_units;
    double cutoff;               // r^2 cut-off in source units
};

// Reaction sums that belong to bodies of OTHER shards leave through nb_sym_push_kernel: the rows
// of gacc that cover peer p's shard go to this rank's slot in p's receive buffer (plain coalesced
// stores over NVLink peer memory), the local rows are cleared, and the last CTA publishes the
// pass number on p's flag word.  The owner's nb_finish_kernel acquires the flags of all its senders.
struct NbSymPush {
    double* gacc;
    size_t gstride;
    int n_dst;
    long long body_begin[4];     // first body of the destination shard
    int count;                   // bodies per shard (incl. tile padding)
    double* dst_slot[4];         // peer-mapped slot [3][count]
    unsigned long long* dst_flag[4];   // peer-mapped flag word of (this rank -> peer)
    unsigned long long seq;
    unsigned* done;              // CTA completion counter (self-resetting)
};

struct NbSymFinish {
    double* gacc;
    size_t gstride;
    int n_src;                   // senders whose slots must be added
    const double* slot[4];       // local receive slots [3][count]
    const unsigned long long* flag[4];   // local flag words of (sender -> this rank)
    int sender[4];               // their ranks (for the timeout report)
    unsigned long long seq;      // 0 = nothing to wait for
    int count;
    unsigned* done;              // CTA completion counter for the step signal (self-resetting)
    float* lo_next;              // 48-bit positions: lo rows of the next step's sources (null otherwise)
};

// TMA ring depth of a shape: the 4 x 128 FP32 shape was sized for five CTAs per SM (two stages); it runs four now
// (one tile takes ~10 us to consume, a bulk copy ~1 us to land: two are enough)
__host__ __device__ constexpr int nb_sym_stages(bool f64, int ti, int block) {
    return ((!f64 && ti == 4 && block == 128) || (f64 && ti == 2)) ? 2 : NB_STAGES;
}
// Resident CTAs per SM the register allocation is capped for.  The 8 x 128 FP32 shape in 3D with per-body masses is
// compiled for TWO CTAs per SM (245 registers, two warps per scheduler): ptxas keeps more chains in flight per warp and
// that beats three CTAs at 168 registers by 5 % (3899 vs 3712 G inter/s at N = 2^20, same arithmetic); the shorter 2D
// and equal-mass chains are faster with three (2D N = 65536: 4627 vs 4423; Plummer N = 262144: 4391 vs 4167).
// 4 x 128 FP32 (small problems): four CTAs per SM at 128 registers.  Five CTAs at 96 registers are 1 % faster when the
// library is alone in the process (N = 16384: 0.1002 vs 0.1012 ms/step) but spill 56 bytes per thread, and the cost of
// that local-memory traffic depends on the rest of the CUDA context: 0.1071 ms once PyTorch has launched one kernel
// of its own (lazily loaded module), against 0.1016 for this spill-free build (tools/c2_probe.py).
#ifndef NB_SYM_MB_4X128
#define NB_SYM_MB_4X128 4
#endif
__host__ __device__ constexpr int nb_sym_min_blocks(bool f64, int ti, int block, int dim = 3, bool eqm = false) {
    return block == 64 ? (f64 ? 4 : 7)
           : f64 ? (ti == 2 ? (block == 256 ? 2 : 4) : ti == 8 ? 2 : block == 128 ? 3 : 1)
                 : block == 256 ? 2 : (ti == 8 ? ((dim == 3 && !eqm) ? 2 : 3) : NB_SYM_MB_4X128);
}

static inline size_t nb_sym_smem_bytes(int dim, int block, bool f64, int ti = 0, int algo = 0, bool hl = false) {
    const size_t rs = f64 ? 8 : 4;
    if (ti == 0) ti = block == 64 ? 4 : NB_SYM_ITILE / block;
    const int stages = nb_sym_stages(f64, ti, block);
    const size_t ring = (size_t)stages * NB_TILE * (dim + 1 + (hl ? dim : 0)) * rs;
    const size_t bars = 2 * stages * sizeof(uint64_t) + 16;
    // transpose scratch of sym_algo 0; the rotation flavours keep the FP64 per-target sums there
    // ([TI * 3][block] doubles)
    const size_t scr = std::max(algo == 0 ? (size_t)(block / 32) * 2 * 32 * NB_SYM_ROW * sizeof(float) : (size_t)0,
                                (f64 || algo == 0) ? (size_t)0 : (size_t)3 * ti * block * sizeof(double));
    const size_t bout = (size_t)2 * (block / 32) * dim * NB_TILE * rs;
    return ring + bars + scr + bout;
}

// One source tile against this thread's four targets, both directions.
//   a[t][d]  += sum_j (m_j / r^4) d_ij          (this tile's FP32 partial of the target sums)
//   wout[d][j] = sum over the warp's 128 targets of (m_i / r^4) d_ij   (NOT yet negated)
// scr = this warp's two 32 x NB_SYM_ROW transpose buffers; lane = tid & 31.
template <int D, int TI, int MODE>
__device__ __forceinline__ void nb_tile_f32_sym(const float* __restrict__ stage, float cutoff,
                                                const float (&npos)[TI][3],
                                                const float (&mi)[TI],
                                                float2 (&a)[TI][3], float* __restrict__ scr,
                                                float* __restrict__ wout, int lane) {
    const float4* sx = reinterpret_cast<const float4*>(stage);
    const float4* sy = sx + NB_TILE / 4;
    const float4* sz = sy + NB_TILE / 4;                      // D == 3 only
    const float4* sm = sx + D * (NB_TILE / 4);
#pragma unroll
    for (int t = 0; t < TI; ++t)
#pragma unroll
        for (int d = 0; d < 3; ++d) a[t][d] = make_float2(0.f, 0.f);
    const float inf = __int_as_float(0x7f800000);

    // column-sum role of this lane in the transpose: output o = component * 4 + source (the word
    // index inside a lane row), half hh = even / odd lane rows.  Word (2k + hh) * 14 + o lies in bank
    // (28 k + 14 hh + o) mod 32: the 24 reducer lanes hit 24 distinct banks for every k.
    const int o = lane >> 1, hh = lane & 1;
    const bool reducer = lane < 8 * D;
    const int oc = reducer ? o : 0;
    const float* rbase = scr + hh * NB_SYM_ROW + oc;
    float* wbase = wout + (oc >> 2) * NB_TILE + (oc & 3);
    const bool writer = reducer && hh == 0;
    auto reduce_store = [&](int q_done, int buf) {
        const float* rb = rbase + buf * (32 * NB_SYM_ROW);
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
            v0 += rb[(2 * k) * NB_SYM_ROW];
            v1 += rb[(2 * k + 2) * NB_SYM_ROW];
            v2 += rb[(2 * k + 4) * NB_SYM_ROW];
            v3 += rb[(2 * k + 6) * NB_SYM_ROW];
        }
        float v = (v0 + v1) + (v2 + v3);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (writer && q_done >= 0) wbase[q_done * 4] = v;
    };

#ifndef NB_SYM_UNROLL
#define NB_SYM_UNROLL 1
#endif
    NB_UNROLL(NB_SYM_UNROLL)
    for (int q = 0; q < NB_TILE / 4; ++q) {
        const float4 X = sx[q], Y = sy[q], M = sm[q];
        float4 Z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (D == 3) Z = sz[q];
        float2 b[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float2 xs = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
            const float2 ys = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
            const float2 zs = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
            const float2 ms = h ? make_float2(M.z, M.w) : make_float2(M.x, M.y);
#pragma unroll
            for (int d = 0; d < 3; ++d) b[h][d] = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < TI; ++t) {
                const float2 dx = __fadd2_rn(xs, make_float2(npos[t][0], npos[t][0]));
                const float2 dy = __fadd2_rn(ys, make_float2(npos[t][1], npos[t][1]));
                float2 r2 = __fmul2_rn(dx, dx);
                r2 = __ffma2_rn(dy, dy, r2);
                float2 dz;
                if (D == 3) {
                    dz = __fadd2_rn(zs, make_float2(npos[t][2], npos[t][2]));
                    r2 = __ffma2_rn(dz, dz, r2);
                }
                if (MODE == NB_EXACT) {
                    r2.x = (r2.x >= cutoff) ? r2.x : inf;
                    r2.y = (r2.y >= cutoff) ? r2.y : inf;
                }
                float2 inv;
                inv.x = nb_rcp_f32(r2.x);
                inv.y = nb_rcp_f32(r2.y);
                const float2 w = __fmul2_rn(inv, inv);
                const float2 s = __fmul2_rn(w, ms);
                const float2 u = __fmul2_rn(w, make_float2(mi[t], mi[t]));
                a[t][0] = __ffma2_rn(dx, s, a[t][0]);
                b[h][0] = __ffma2_rn(dx, u, b[h][0]);
                a[t][1] = __ffma2_rn(dy, s, a[t][1]);
                b[h][1] = __ffma2_rn(dy, u, b[h][1]);
                if (D == 3) {
                    a[t][2] = __ffma2_rn(dz, s, a[t][2]);
                    b[h][2] = __ffma2_rn(dz, u, b[h][2]);
                }
            }
        }
        // the previous iteration's rows are complete (syncwarp at its end): reduce them while
        // this iteration's chains are in flight, then publish this iteration's rows
        reduce_store(q - 1, (q + 1) & 1);   // q = 0: nothing stored
        float2* row = reinterpret_cast<float2*>(scr + (q & 1) * (32 * NB_SYM_ROW) + lane * NB_SYM_ROW);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            row[2 * d] = b[0][d];
            row[2 * d + 1] = b[1][d];
        }
        __syncwarp();
    }
    reduce_store(NB_TILE / 4 - 1, (NB_TILE / 4 - 1) & 1);
    __syncwarp();
}

// The same tile, reaction sums kept in REGISTERS and rotated through the warp (ALGO 1 / 2).
// A half tile = 128 sources = 32 "home groups" of four consecutive sources, home group h living in
// lane h.  In step k lane l works on home group (l + k) & 31: it reads the group's coordinates from
// the shared-memory stage (one LDS.128 per plane, the lanes hit 32 different 16-byte chunks: conflict
// free), adds its TI targets' reactions to the group's travelling partial sums and hands those to lane
// l - 1, which meets the group next (12 SHFL per 8 chains).  After 32 steps every group's sums are
// back in their home lane, complete for the warp's 128 targets x 128 sources, and leave through one
// STS.128 per component.  Against the transpose flavour above this drops, per 8 chains, 24 LDS +
// 7 STS + 18 scalar FADD + the __syncwarp and keeps 12 SHFL.
//   DECOUPLE = false: the chains accumulate straight onto the travelling sums; each pair of sources
//                     is shuffled as soon as its TI chains are done, so the other pair's chains cover it
//   DECOUPLE = true : per-step local sums; sent = received + local (6 FADD2 more per step); what a
//                     lane receives is first needed a whole step later
//   EQM = true      : equal-mass system (every real body has the same mass, padding bodies are parked out of range):
//                     s = u = 1/r^4, the two mass multiplies leave the chain (13 instead of 15 packed instructions
//                     per two pairs) and the common mass is applied once per body by the finish kernel
//   HL = true       : 48-bit positions (option "fp32_positions" = 48): every coordinate is a float pair hi + lo; the
//                     difference is (hi_j - hi_i) + (lo_j - lo_i) -- the first term is exact for close pairs, so the
//                     24-bit quantisation of the positions no longer moves the near field (2 FADD2 more per coordinate:
//                     21 instead of 15 packed instructions per two pairs).  The lo planes follow the D + 1 hi planes
//                     in the stage; nlo holds the negated lo parts of the targets.
template <int D, int TI, int MODE, bool DECOUPLE, bool EQM = false, bool HL = false>
__device__ __forceinline__ void nb_tile_f32_sym_rot(const float* __restrict__ stage, float cutoff,
                                                    const float (&npos)[TI][3],
                                                    const float (&mi)[TI],
                                                    float2 (&a)[TI][3], float* __restrict__ wout, int lane,
                                                    int hf_begin, int hf_end, const float (*nlo)[3] = nullptr) {
    // plane p of the stage as float4: home group g of half tile hf sits at p * (NB_TILE / 4) + 32 hf + g
    const float4* sx = reinterpret_cast<const float4*>(stage);
    const float4* sy = sx + NB_TILE / 4;
    const float4* sz = sy + NB_TILE / 4;                      // D == 3 only
    const float4* sm = sx + D * (NB_TILE / 4);
    const float4* lx = sm + NB_TILE / 4;                      // HL only: lo planes
    const float4* ly = lx + NB_TILE / 4;
    const float4* lz = ly + NB_TILE / 4;
#pragma unroll
    for (int t = 0; t < TI; ++t)
#pragma unroll
        for (int d = 0; d < 3; ++d) a[t][d] = make_float2(0.f, 0.f);
    const float inf = __int_as_float(0x7f800000);
    const int from = (lane + 1) & 31;

    // the TI chains of one pair of sources: target sums a, reaction sums b
    auto chains = [&](const float2 xs, const float2 ys, const float2 zs, const float2 ms, float2 (&b)[3],
                      const float2 xl, const float2 yl, const float2 zl) {
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
            }
            float2 inv;
            inv.x = nb_rcp_f32(r2.x);
            inv.y = nb_rcp_f32(r2.y);
            const float2 w = __fmul2_rn(inv, inv);
            const float2 s = EQM ? w : __fmul2_rn(w, ms);
            const float2 u = EQM ? w : __fmul2_rn(w, make_float2(mi[t], mi[t]));
            a[t][0] = __ffma2_rn(dx, s, a[t][0]);
            b[0] = __ffma2_rn(dx, u, b[0]);
            a[t][1] = __ffma2_rn(dy, s, a[t][1]);
            b[1] = __ffma2_rn(dy, u, b[1]);
            if (D == 3) {
                a[t][2] = __ffma2_rn(dz, s, a[t][2]);
                b[2] = __ffma2_rn(dz, u, b[2]);
            }
        }
    };
    // this pair of sources is done for this lane: pass its sums on to lane l - 1
    auto hand_over = [&](float2 (&trav)[3], const float2 (&b)[3]) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float2 v = DECOUPLE ? __fadd2_rn(trav[d], b[d]) : b[d];
            // in-place shuffles: the loop-carried pair keeps its registers (no MOVs at the loop end)
            asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+f"(v.x) : "r"(from));
            asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+f"(v.y) : "r"(from));
            trav[d] = v;
        }
    };
    const float2 zero2 = make_float2(0.f, 0.f);

#pragma unroll 1
    for (int hf = hf_begin; hf < hf_end; ++hf) {
        float2 trav[2][3];                                    // travelling sums of the group this lane meets next
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int d = 0; d < 3; ++d) trav[h][d] = zero2;
#pragma unroll 1
        for (int k = 0; k < 32; ++k) {
            const int q = hf * 32 + ((lane + k) & 31);
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
                float2 b[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) b[d] = DECOUPLE ? zero2 : trav[h][d];
                chains(xs, ys, zs, ms, b, xl, yl, zl);
                hand_over(trav[h], b);
            }
        }
        // after 32 hand-overs the sums of home group `lane` are complete and back in lane `lane`
        float4* wo = reinterpret_cast<float4*>(wout + hf * 128 + 4 * lane);
#pragma unroll
        for (int d = 0; d < D; ++d)
            wo[d * (NB_TILE / 4)] = make_float4(trav[0][d].x, trav[0][d].y, trav[1][d].x, trav[1][d].y);
    }
}

// FP64 flavour: scalar DFMA chains, two sources per iteration, sums kept in FP64 end to end.
// Per unordered pair 18 DP operations (3 DADD, DMUL + 2 DFMA, 3 DFMA of the reciprocal, DMUL for
// 1/r^4, 2 DMUL for the two masses, 2 x 3 DFMA) instead of 2 x 14.  The lane rows of the transpose hold
// 6 doubles (stride 7): reducer lane (o, part) adds the rows 4k + part of output o = component*2 + source.
template <int D, int TI, bool EXACT>
__device__ __forceinline__ void nb_tile_f64_sym(const double* __restrict__ stage, double cutoff,
                                                const double (&pos)[TI][3], const double (&mi)[TI],
                                                double (&accd)[TI][3], double* __restrict__ scr,
                                                double* __restrict__ wout, int lane) {
    constexpr int ROWD = NB_SYM_ROW / 2;                      // doubles per lane row
    const double2* sx = reinterpret_cast<const double2*>(stage);
    const double2* sy = sx + NB_TILE / 2;
    const double2* sz = sy + NB_TILE / 2;                     // D == 3 only
    const double2* sm = sx + D * (NB_TILE / 2);
    const int o = lane >> 2, part = lane & 3;
    const bool reducer = o < 2 * D;
    const int oc = reducer ? o : 0;
    const double* rbase = scr + part * ROWD + oc;
    double* wbase = wout + (oc >> 1) * NB_TILE + (oc & 1);
    const bool writer = reducer && part == 0;
    auto reduce_store = [&](int q_done, int buf) {
        const double* rb = rbase + buf * (32 * ROWD);
        double v0 = 0.0, v1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            v0 += rb[(4 * k) * ROWD];
            v1 += rb[(4 * k + 4) * ROWD];
        }
        double v = v0 + v1;
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (writer && q_done >= 0) wbase[q_done * 2] = v;
    };

#pragma unroll 1
    for (int q = 0; q < NB_TILE / 2; ++q) {
        const double2 X = sx[q], Y = sy[q], M = sm[q];
        double2 Z = make_double2(0.0, 0.0);
        if (D == 3) Z = sz[q];
        double b[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double xs = h ? X.y : X.x, ys = h ? Y.y : Y.x, zs = h ? Z.y : Z.x;
            const double ms = h ? M.y : M.x;
#pragma unroll
            for (int d = 0; d < 3; ++d) b[h][d] = 0.0;
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
                const double w = inv * inv;
                const double s = w * ms;
                const double u = w * mi[t];
                accd[t][0] = fma(dx, s, accd[t][0]);
                b[h][0] = fma(dx, u, b[h][0]);
                accd[t][1] = fma(dy, s, accd[t][1]);
                b[h][1] = fma(dy, u, b[h][1]);
                if (D == 3) {
                    accd[t][2] = fma(dz, s, accd[t][2]);
                    b[h][2] = fma(dz, u, b[h][2]);
                }
            }
        }
        reduce_store(q - 1, (q + 1) & 1);   // q = 0: nothing stored
        double* row = scr + (q & 1) * (32 * ROWD) + lane * ROWD;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            row[2 * d] = b[0][d];
            row[2 * d + 1] = b[1][d];
        }
        __syncwarp();
    }
    reduce_store(NB_TILE / 2 - 1, (NB_TILE / 2 - 1) & 1);
    __syncwarp();
}

// FP64, reaction sums rotated through the warp (see nb_tile_f32_sym_rot): a quarter tile = 64 sources =
// 32 home groups of two, one LDS.128 per plane and step, 12 SHFL (six doubles) per 2 x TI chains.
template <int D, int TI, bool EXACT, bool EQM = false>
__device__ __forceinline__ void nb_tile_f64_sym_rot(const double* __restrict__ stage, double cutoff,
                                                    const double (&pos)[TI][3], const double (&mi)[TI],
                                                    double (&accd)[TI][3], double* __restrict__ wout, int lane,
                                                    int qd_begin, int qd_end) {
    const double2* sx = reinterpret_cast<const double2*>(stage);
    const double2* sy = sx + NB_TILE / 2;
    const double2* sz = sy + NB_TILE / 2;                     // D == 3 only
    const double2* sm = sx + D * (NB_TILE / 2);
    const int from = (lane + 1) & 31;
#pragma unroll 1
    for (int qd = qd_begin; qd < qd_end; ++qd) {
        double trav[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int d = 0; d < 3; ++d) trav[h][d] = 0.0;
#pragma unroll 1
        for (int k = 0; k < 32; ++k) {
            const int g = qd * 32 + ((lane + k) & 31);
            const double2 X = sx[g], Y = sy[g], M = sm[g];
            double2 Z = make_double2(0.0, 0.0);
            if (D == 3) Z = sz[g];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double xs = h ? X.y : X.x, ys = h ? Y.y : Y.x, zs = h ? Z.y : Z.x;
                const double ms = h ? M.y : M.x;
                double b[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) b[d] = trav[h][d];
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
                    const double w = inv * inv;
                    const double s = EQM ? w : w * ms;
                    const double u = EQM ? w : w * mi[t];
                    accd[t][0] = fma(dx, s, accd[t][0]);
                    b[0] = fma(dx, u, b[0]);
                    accd[t][1] = fma(dy, s, accd[t][1]);
                    b[1] = fma(dy, u, b[1]);
                    if (D == 3) {
                        accd[t][2] = fma(dz, s, accd[t][2]);
                        b[2] = fma(dz, u, b[2]);
                    }
                }
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    int lo = __double2loint(b[d]), hi = __double2hiint(b[d]);
                    asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(lo) : "r"(from));
                    asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(hi) : "r"(from));
                    trav[h][d] = __hiloint2double(hi, lo);
                }
            }
        }
        double2* wo = reinterpret_cast<double2*>(wout + qd * 64 + 2 * lane);
#pragma unroll
        for (int d = 0; d < D; ++d) wo[d * (NB_TILE / 2)] = make_double2(trav[0][d], trav[1][d]);
    }
}

// ALGO: 0 = shared-memory transpose of the reaction sums, 1 = register rotation, 2 (FP32) = rotation, decoupled
//       HL (FP32 rotation, general masses): 48-bit positions -- a stage holds the D + 1 hi planes of a tile followed by
//       its D lo planes (two bulk copies on one barrier); the pre-pass, which sees the hi parts, widens its cells by the
//       largest hi-part difference error so that its flags stay conservative
template <int D, bool F64, int TI, int BLOCK, int ALGO = 0, bool EQM = false, bool HL = false>
__global__ void __launch_bounds__(BLOCK, nb_sym_min_blocks(F64, TI, BLOCK, D, EQM))
nb_force_sym_kernel(const NbSymParams P) {
    static_assert(!EQM || ALGO != 0, "the equal-mass flavour exists for the rotation flavours only");
    static_assert(!HL || (!F64 && ALGO == 2 && !EQM), "48-bit positions: FP32, decoupled rotation, general masses");
    constexpr int STAGES = nb_sym_stages(F64, TI, BLOCK);
    using real = typename NbReal<F64>::type;
    constexpr int NP = D + 1;
    constexpr int ITILE = TI * BLOCK;
    static_assert(ITILE % NB_TILE == 0, "an i-tile is a whole number of source tiles");
    constexpr int TILE_ELEMS = NB_TILE * NP;
    constexpr uint32_t TILE_BYTES = TILE_ELEMS * sizeof(real);
    constexpr int LO_ELEMS = HL ? NB_TILE * D : 0;
    constexpr uint32_t LO_BYTES = LO_ELEMS * sizeof(real);
    constexpr int STAGE_ELEMS = TILE_ELEMS + LO_ELEMS;
    constexpr int NWARPS = BLOCK / 32;

    extern __shared__ __align__(128) unsigned char nb_smem[];
    real* ring = reinterpret_cast<real*>(nb_smem);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(nb_smem + (size_t)STAGES * STAGE_ELEMS * sizeof(real));
    uint64_t* empty_bar = full_bar + STAGES;
    int* s_unit = reinterpret_cast<int*>(empty_bar + STAGES);   // [0] i-tile (or -1), [1] segment
    float* scr_all = reinterpret_cast<float*>(s_unit + 4);
    // ALGO 0: the transpose scratch; FP32 rotation: the FP64 per-target sums; FP64 rotation: nothing
    constexpr size_t SCR_FLOATS = ALGO == 0 ? (size_t)NWARPS * 2 * 32 * NB_SYM_ROW : F64 ? (size_t)0 : (size_t)6 * ITILE;
    real* bout_all = reinterpret_cast<real*>(scr_all + SCR_FLOATS);   // [2][NWARPS][D][NB_TILE]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const real* __restrict__ src = static_cast<const real*>(P.src);
    float* scr = scr_all + (size_t)warp * 2 * 32 * NB_SYM_ROW;

    nb_launch_dependents();   // the finish kernel behind this pass may queue up (it waits for the pass's completion itself)
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            nb_mbar_init(&full_bar[s], 1);
            nb_mbar_init(&empty_bar[s], NWARPS);
        }
        nb_fence_mbar_init();
    }
    __syncthreads();
    nb_grid_dep_wait();       // sources, accumulators and the unit counter come from the kernels before this one

    unsigned kt = 0;          // tiles consumed by this CTA so far (ring position)
    int bbuf = 0;             // which half of bout the next symmetric tile writes

    for (;;) {
        if (tid == 0) {
            // (fetching the next unit's number one unit ahead was measured: 2 % slower at N = 2^20, 30 % at N = 16384)
            const int u = (int)atomicAdd(&P.sched[0], 1u);
            int r = -1, sg = 0;
            if (u < P.total_units) {
                const int* prefix = P.row_prefix;
                int lo = 0, hi = P.n_rows;              // largest row with row_prefix[row] <= u
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (prefix[mid] <= u) lo = mid; else hi = mid;
                }
                r = lo;
                sg = u - prefix[lo];
            }
            s_unit[0] = r;
            s_unit[1] = sg;
        }
        __syncthreads();
        const int row = s_unit[0];
        const int sg = s_unit[1];
        __syncthreads();
        if (row < 0) break;
        const NbSymRow R = P.rows[row];
        const int it = R.it;
        const bool sym = (R.flags & NB_ROW_SYM) != 0;
        // a symmetric row is cut in units of seg_sub sub-tiles, an ordered row in units of seg_ord tiles
        constexpr int SUBT = nb_sym_subtiles(F64, ALGO);
        const int rsub = sym ? SUBT : 1;
        const int s0 = R.t_begin * rsub + sg * (sym ? P.seg_sub : P.seg_ord);
        const int s1 = min(s0 + (sym ? P.seg_sub : P.seg_ord), R.t_end * rsub);
        const int ts = s0 / rsub;
        const int te = (s1 + rsub - 1) / rsub;
        const int ntl = te - ts;

        if (tid == 0) {
            const int pre = min(STAGES - 1, ntl);
            for (int t = 0; t < pre; ++t) {
                const unsigned k = kt + t;
                const int slot = k % STAGES;
                nb_mbar_wait(&empty_bar[slot], ((k / STAGES) & 1u) ^ 1u);
                nb_mbar_expect_tx(&full_bar[slot], TILE_BYTES + LO_BYTES);
                nb_tma_load_1d(ring + (size_t)slot * STAGE_ELEMS, src + (size_t)(ts + t) * TILE_ELEMS,
                               TILE_BYTES, &full_bar[slot]);
                if constexpr (HL)
                    nb_tma_load_1d(ring + (size_t)slot * STAGE_ELEMS + TILE_ELEMS,
                                   reinterpret_cast<const real*>(P.src_lo) + (size_t)(ts + t) * LO_ELEMS, LO_BYTES, &full_bar[slot]);
            }
        }

        real tq[TI][3], mi[TI];          // FP32: NEGATED target coordinates (operand of the packed add); FP64: plain
        float tlo[HL ? TI : 1][3];       // HL: negated lo parts
        int own_tile[TI];
        bool suspect = false;
#pragma unroll
        for (int t = 0; t < TI; ++t) {
            const long long b = P.tgt_base + (long long)it * ITILE + tid + t * BLOCK;
            own_tile[t] = (int)(b / NB_TILE);
            const real* tb = src + (size_t)(b / NB_TILE) * TILE_ELEMS + (b % NB_TILE);
            // The last i-tile of a shard may reach past the shard's bodies (into the next shard's, which
            // also appear as SOURCES of the cross-shard rows: a self pair outside the exact pass).  Such
            // lanes are made inert: parked far away (1/r^4 underflows to 0) with zero mass, nothing stored.
            const bool live = it * ITILE + tid + t * BLOCK < P.own_count;
            const real park = F64 ? real(1.0e150) : real(1.0e15f);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const real x = (d < D) ? (live ? tb[d * NB_TILE] : park) : real(0);
                tq[t][d] = F64 ? x : -x;
                if constexpr (HL) {
                    const float* tl = P.src_lo + (size_t)(b / NB_TILE) * LO_ELEMS + (b % NB_TILE);
                    tlo[t][d] = (d < D && live) ? -tl[d * NB_TILE] : 0.f;
                }
            }
            mi[t] = live ? tb[D * NB_TILE] : real(0);
            if (P.suspect) suspect |= P.suspect[it * ITILE + tid + t * BLOCK] != 0;
        }
        const bool warp_suspect = !P.suspect || __any_sync(0xffffffffu, suspect) != 0;
        // FP64 sums of the own targets over the unit: in registers, or (FP32 rotation flavours) in the shared memory
        // the transpose scratch no longer needs -- 24 registers back for the chains
        constexpr bool SACC = !F64 && ALGO != 0;
        double* sacc = reinterpret_cast<double*>(scr_all) + tid;     // [TI * 3][BLOCK]
        double accd[SACC ? 1 : TI][3];
        if constexpr (SACC) {
#pragma unroll
            for (int k = 0; k < TI * 3; ++k) sacc[k * BLOCK] = 0.0;
        } else {
#pragma unroll
            for (int t = 0; t < TI; ++t)
#pragma unroll
                for (int d = 0; d < 3; ++d) accd[t][d] = 0.0;
        }

        for (int t = 0; t < ntl; ++t) {
            if (tid == 0 && t + STAGES - 1 < ntl) {
                const unsigned k = kt + t + STAGES - 1;
                const int slot = k % STAGES;
                nb_mbar_wait(&empty_bar[slot], ((k / STAGES) & 1u) ^ 1u);
                nb_mbar_expect_tx(&full_bar[slot], TILE_BYTES + LO_BYTES);
                nb_tma_load_1d(ring + (size_t)slot * STAGE_ELEMS,
                               src + (size_t)(ts + t + STAGES - 1) * TILE_ELEMS, TILE_BYTES,
                               &full_bar[slot]);
                if constexpr (HL)
                    nb_tma_load_1d(ring + (size_t)slot * STAGE_ELEMS + TILE_ELEMS,
                                   reinterpret_cast<const real*>(P.src_lo) + (size_t)(ts + t + STAGES - 1) * LO_ELEMS, LO_BYTES,
                                   &full_bar[slot]);
            }
            const unsigned k = kt + t;
            const int slot = k % STAGES;
            nb_mbar_wait(&full_bar[slot], (k / STAGES) & 1u);
            const real* stage = ring + (size_t)slot * STAGE_ELEMS;
            real* wout = bout_all + ((size_t)bbuf * NWARPS + warp) * (D * NB_TILE);
            // sub-tiles [h0, h1) of this tile belong to the unit (whole tile: [0, SUBT))
            const int h0 = sym ? max(s0 - (ts + t) * SUBT, 0) : 0;
            const int h1 = sym ? min(s1 - (ts + t) * SUBT, SUBT) : SUBT;
            bool exact_tile = warp_suspect;
            if (!sym) {
#pragma unroll
                for (int tt = 0; tt < TI; ++tt) exact_tile |= (ts + t == own_tile[tt]);
            }
            if constexpr (F64) {
                static_assert(!SACC, "shared-memory accumulators belong to the FP32 rotation flavours");
                const double* dstage = reinterpret_cast<const double*>(stage);
                const double(&pos)[TI][3] = reinterpret_cast<const double(&)[TI][3]>(tq);
                const double(&mid)[TI] = reinterpret_cast<const double(&)[TI]>(mi);
                double* dscr = reinterpret_cast<double*>(scr);
                double* dwout = reinterpret_cast<double*>(wout);
                if (sym) {
                    if constexpr (ALGO == 0) {
                        if (exact_tile) nb_tile_f64_sym<D, TI, true>(dstage, P.cutoff, pos, mid, accd, dscr, dwout, lane);
                        else nb_tile_f64_sym<D, TI, false>(dstage, P.cutoff, pos, mid, accd, dscr, dwout, lane);
                    } else {
                        if (exact_tile) nb_tile_f64_sym_rot<D, TI, true, EQM>(dstage, P.cutoff, pos, mid, accd, dwout, lane, h0, h1);
                        else nb_tile_f64_sym_rot<D, TI, false, EQM>(dstage, P.cutoff, pos, mid, accd, dwout, lane, h0, h1);
                    }
                } else {
                    if (exact_tile) nb_tile_f64<D, TI, 1, true, EQM>(dstage, 0, P.cutoff, pos, accd);
                    else nb_tile_f64<D, TI, 1, false, EQM>(dstage, 0, P.cutoff, pos, accd);
                }
            } else {
                const float* fstage = reinterpret_cast<const float*>(stage);
                const float(&npos)[TI][3] = reinterpret_cast<const float(&)[TI][3]>(tq);
                const float(&mif)[TI] = reinterpret_cast<const float(&)[TI]>(mi);
                const float cutoff_f = (float)P.cutoff;
                float2 a[TI][3];
                if (sym) {
                    float* fwout = reinterpret_cast<float*>(wout);
                    if constexpr (ALGO == 0) {
                        if (exact_tile) nb_tile_f32_sym<D, TI, NB_EXACT>(fstage, cutoff_f, npos, mif, a, scr, fwout, lane);
                        else nb_tile_f32_sym<D, TI, NB_PLAIN>(fstage, cutoff_f, npos, mif, a, scr, fwout, lane);
                    } else {
                        if constexpr (HL) {
                            if (exact_tile) nb_tile_f32_sym_rot<D, TI, NB_EXACT, true, false, true>(fstage, cutoff_f, npos, mif, a, fwout, lane, h0, h1, tlo);
                            else nb_tile_f32_sym_rot<D, TI, NB_PLAIN, true, false, true>(fstage, cutoff_f, npos, mif, a, fwout, lane, h0, h1, tlo);
                        } else if (exact_tile) nb_tile_f32_sym_rot<D, TI, NB_EXACT, ALGO == 2, EQM>(fstage, cutoff_f, npos, mif, a, fwout, lane, h0, h1);
                        else nb_tile_f32_sym_rot<D, TI, NB_PLAIN, ALGO == 2, EQM>(fstage, cutoff_f, npos, mif, a, fwout, lane, h0, h1);
                    }
                } else {
                    if constexpr (HL) {
                        if (exact_tile) nb_tile_f32<D, TI, 1, NB_EXACT, 1, false, true>(fstage, 0, cutoff_f, npos, a, tlo);
                        else nb_tile_f32<D, TI, 1, NB_PLAIN, 1, false, true>(fstage, 0, cutoff_f, npos, a, tlo);
                    } else if (exact_tile) nb_tile_f32<D, TI, 1, NB_EXACT, 1, EQM>(fstage, 0, cutoff_f, npos, a);
                    else nb_tile_f32<D, TI, 1, NB_PLAIN, 1, EQM>(fstage, 0, cutoff_f, npos, a);
                }
#pragma unroll
                for (int tt = 0; tt < TI; ++tt)
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        if constexpr (SACC) sacc[(tt * 3 + d) * BLOCK] += (double)(a[tt][d].x + a[tt][d].y);
                        else accd[tt][d] += (double)(a[tt][d].x + a[tt][d].y);
                    }
            }
            __syncwarp();
            if (lane == 0) nb_mbar_arrive(&empty_bar[slot]);
            if (sym) {
                // CTA-wide sum of the eight warps' tile buffers, one FP64 atomic per (source, component);
                // the reaction on source j is MINUS sum_i (m_i / r^4) d_ij
                __syncthreads();
                const real* bb = bout_all + (size_t)bbuf * NWARPS * (D * NB_TILE);
                for (int j = h0 * (NB_TILE / SUBT) + tid; j < h1 * (NB_TILE / SUBT); j += BLOCK) {
                    const size_t gj = (size_t)(ts + t) * NB_TILE + j;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        real v = real(0);     // (summing the warps' FP32 partials in FP64 was measured: 2-7 % slower, no gain in accuracy)
#pragma unroll
                        for (int w = 0; w < NWARPS; ++w) v += bb[(size_t)w * (D * NB_TILE) + d * NB_TILE + j];
                        atomicAdd(&P.gacc[(size_t)d * P.gstride + gj], -(double)v);
                    }
                }
                bbuf ^= 1;
            }
        }
        kt += ntl;

#pragma unroll
        for (int t = 0; t < TI; ++t)
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const int li = it * ITILE + tid + t * BLOCK;
                if (li < P.own_count)
                    atomicAdd(&P.gacc[(size_t)d * P.gstride + (size_t)P.tgt_base + li],
                              SACC ? sacc[(t * 3 + d) * BLOCK] : accd[SACC ? 0 : t][d]);
            }
    }

    if (tid == 0) {
        __threadfence();
        const unsigned e = atomicAdd(&P.sched[1], 1u);
        if (e + 1u == gridDim.x) {
            P.sched[0] = 0u;
            P.sched[1] = 0u;
            __threadfence();
        }
    }
}

template <int D>
__global__ void __launch_bounds__(256) nb_sym_push_kernel(const NbSymPush Q) {
    const long long per = (long long)Q.count * D;
    const long long total = per * Q.n_dst;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total;
         k += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(k / per);
        const long long r = k - (long long)p * per;
        const int d = (int)(r / Q.count);
        const long long j = r - (long long)d * Q.count;
        double* g = &Q.gacc[(size_t)d * Q.gstride + (size_t)(Q.body_begin[p] + j)];
        Q.dst_slot[p][(size_t)d * Q.count + j] = *g;
        *g = 0.0;
    }
    __threadfence_system();                       // this thread's peer stores before the CTA's exit count
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned e = atomicAdd(Q.done, 1u);
        if (e + 1u == gridDim.x) {
            *Q.done = 0u;
            __threadfence_system();
            for (int p = 0; p < Q.n_dst; ++p) nb_st_release_sys(Q.dst_flag[p], Q.seq);
        }
    }
}

// Epilogue of a pass whose accumulators are complete only when the whole pass is (symmetric
// pass): the same arithmetic as the fused epilogue of nb_force_kernel, after adding the reaction
// sums the other ranks pushed into this rank's receive slots.
//   forces:  F_i = -(G m_i) S_i                                  methods.cpp:125-131
//   step:    v += (F/m) dt ; x += v dt ; new source row           methods.cpp:436, :448
// In step mode the new row also goes to every peer's next buffer (the fused exchange) and the last
// CTA publishes the step number on the peers' step flags, exactly like nb_force_kernel's exit.
template <int D, typename real>
__global__ void __launch_bounds__(256) nb_finish_kernel(const NbForceParams P, const NbSymFinish F) {
    constexpr int NP = D + 1;
    nb_grid_dep_wait();       // the pass before this kernel is complete and visible
    nb_launch_dependents();   // the next step's pass may queue up behind this kernel (it waits for it itself)
    if (F.seq != 0ull && F.n_src > 0) {
        if (threadIdx.x < F.n_src)
            nb_wait_flag(F.flag[threadIdx.x], F.seq, P.spin_timeout_ns, P.err_word, NB_WAIT_REACTION, F.sender[threadIdx.x]);
        __syncthreads();
    }
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li < P.tpad) {
        double S[3];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double* ap = &F.gacc[(size_t)d * F.gstride + (size_t)P.tgt_base + li];
            double v = *ap;
            *ap = 0.0;                                   // self-clean for the next pass
            if (li < F.count)
                for (int k = 0; k < F.n_src; ++k) v += F.slot[k][(size_t)d * F.count + li];
            S[d] = v * P.acc_scale;
        }
        if (li < P.n_local) {
            const double m = P.mass[li];
            const double gm = P.G * m;
            double Fo[3];
#pragma unroll
            for (int d = 0; d < D; ++d) Fo[d] = -(gm * S[d]);
            if (P.mode == 0) {
#pragma unroll
                for (int d = 0; d < D; ++d) P.forces[(size_t)li * D + d] = Fo[d];
            } else {
                const long long b = P.tgt_base + li;
                const size_t off0 = (size_t)(b / NB_TILE) * (NB_TILE * NP) + (b % NB_TILE);
                real* nb = static_cast<real*>(P.src_next) + off0;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    double v = P.vel[(size_t)d * P.tpad + li];
                    double x = P.pos[(size_t)d * P.tpad + li];
                    v += (Fo[d] / m) * P.dt;
                    x += v * P.dt;
                    P.vel[(size_t)d * P.tpad + li] = v;
                    P.pos[(size_t)d * P.tpad + li] = x;
                    const real xs = (real)(x * P.pos_scale);
                    nb[d * NB_TILE] = xs;
                    if (F.lo_next)
                        F.lo_next[(size_t)(b / NB_TILE) * (NB_TILE * D) + (size_t)d * NB_TILE + (b % NB_TILE)] =
                            (float)(x * P.pos_scale - (double)xs);
                    for (int pr = 0; pr < P.n_peers; ++pr)
                        static_cast<real*>(P.peer_next[pr])[off0 + (size_t)d * NB_TILE] = xs;
                }
            }
        }
    }
    if (P.signal_step != 0ull && P.n_peers > 0) {
        __threadfence_system();                   // remote rows before the exit count
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned e = atomicAdd(F.done, 1u);
            if (e + 1u == gridDim.x) {
                *F.done = 0u;
                __threadfence_system();
                for (int pr = 0; pr < P.n_peers; ++pr)
                    nb_st_release_sys(P.peer_flags[pr] + P.my_rank, P.signal_step);
            }
        }
    }
}
