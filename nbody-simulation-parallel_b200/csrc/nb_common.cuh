// nb_common.cuh -- shared definitions for the B200 brute-force N-body kernels.
//
// Data layout in HBM ("tile-planar" SoA): bodies are grouped in tiles of NB_TILE; inside a
// tile the D coordinates and the mass are separate contiguous planes:
//     [tile t] = x[NB_TILE] | y[NB_TILE] | (z[NB_TILE]) | m[NB_TILE]        (NP = D+1 planes)
// so that (a) one tile is ONE contiguous block (a single cp.async.bulk / TMA copy, and a
// rank's shard is one contiguous all-gather chunk), and (b) LDS.128 of a plane yields four
// consecutive sources = two packed f32x2 operands with no shuffling.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NB_TILE 256          // sources per tile (one TMA bulk copy)
#define NB_STAGES 3          // TMA ring depth
#define NB_MAX_PEERS 7       // other GPUs of one NVSwitch box whose source buffers a shard writes

struct NbForceParams {
    const void* src;         // current tile-planar sources, all npad bodies (float or double)
    void* src_next;          // next-step tile-planar sources (step mode; own shard rows written)
    double* acc;             // [3][tpad] FP64 partial-sum accumulators of the own targets
    unsigned* tile_done;     // [n_itiles] units finished per i-tile (self-resetting)
    unsigned* sched;         // [2] dynamic unit counter + exit counter (self-resetting)
    double* pos;             // [D][tpad] master positions of own targets (FP64)
    double* vel;             // [D][tpad] master velocities
    const double* mass;      // [tpad]    master masses
    double* forces;          // [n_local][D] AoS Vector<D> output (forces mode)
    const unsigned char* suspect;   // [tpad] close-pair flags of the own targets (FLAGS kernels), else null
    long long tgt_base;      // global body index of own target 0 (multiple of NB_TILE)
    long long n_local;       // real (unpadded) own targets
    int tpad;                // padded own targets (multiple of the i-tile)
    int n_itiles;            // tpad / ITILE
    int seg_tiles;           // source tiles per unit
    int rng_begin[3], rng_end[3];   // up to three source-tile ranges of this launch (range 0 first)
    int rng_nseg[3];                // segments (work units per i-tile) in each range
    int lazy_wait;                  // 1: range 0 holds only OWN rows; take the peer handshake when the
                                    //    CTA first reaches a unit of range 1/2 (hides inter-GPU skew)
    unsigned units_per_itile;  // units that must finish (over ALL launches of the step) per i-tile
    int mode;                // 0 = forces, 1 = step (integrate in the epilogue)
    double G;                // gravitational constant (utils.h:21 in the reference)
    double cutoff;           // r^2 cut-off in SOURCE units (scaled for the FP32 path)
    double dt;
    double acc_scale;        // S = acc * acc_scale      (undo the FP32 power-of-two scaling)
    double pos_scale;        // source = pos * pos_scale (FP32 path; 1.0 for FP64)
    // ---- fused position exchange over NVLink peer memory (multi-GPU step mode)
    // The integrator epilogue stores the new source rows of the own targets straight into every
    // peer's NEXT buffer (P2P stores), so no separate all-gather runs.  Step completion is
    // published with one flag per (writer, reader) pair: the last CTA of the pass that ran the
    // epilogues release-stores signal_step into peer_flags[p][my_rank]; a pass that reads remote
    // rows first acquires my_flags[r] >= wait_step for every peer r.
    void* peer_next[NB_MAX_PEERS];                    // peers' next-step source buffers (peer-mapped)
    unsigned long long* peer_flags[NB_MAX_PEERS];     // peers' flag arrays (peer-mapped), indexed by writer rank
    unsigned long long* my_flags;                     // this shard's flag array, written by the peers
    int peer_rank[NB_MAX_PEERS];
    int n_peers;
    int my_rank;
    unsigned long long wait_step;                     // 0 = this pass reads no remote rows newer than the upload
    unsigned long long wait_epoch;                    // peers must have finished this many uploads (repacks)
    unsigned long long signal_step;                   // 0 = this pass publishes nothing
    int flag_stride;                                  // epoch flags live at my_flags[flag_stride + rank]
    // every wait on a peer's flag is bounded: after spin_timeout_ns the waiter records WHAT it waited for in
    // *err_word (page-locked host memory the host reads after the stream sync) and moves on, so a crashed or
    // diverged peer costs an error return (NB200_ESTATE), not a hung GPU
    unsigned long long* err_word;
    unsigned long long spin_timeout_ns;
    // "deterministic" option: instead of meeting in FP64 atomics (whose order varies from run to run) the unit partial
    // sums go to slots[segment][3][tpad] and the epilogue adds them in segment order: bit-identical results from run to
    // run and, with the segments cut on global tile boundaries, between 1 and N GPUs
    double* slots;
    int nseg_total;
};

// what a timed-out wait records: kind << 56 | peer rank << 48 | awaited value (low 48 bits)
#define NB_WAIT_STEP 1ull
#define NB_WAIT_EPOCH 2ull
#define NB_WAIT_REACTION 3ull
#define NB_WAIT_READY 4ull

// ------------------------------------------------------------------ mbarrier / TMA (sm_90+ PTX)
__device__ __forceinline__ uint32_t nb_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void nb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void nb_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void nb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nb_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void nb_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(nb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void nb_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "NB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra NB_DONE_%=;\n"
        "bra NB_WAIT_%=;\n"
        "NB_DONE_%=:\n"
        "}\n" ::"r"(nb_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void nb_tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                               uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(nb_smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(nb_smem_u32(bar))
        : "memory");
}

// system-scope flag accessors for the cross-GPU step handshake
__device__ __forceinline__ unsigned long long nb_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void nb_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may
// start while its predecessor in the stream still runs; nb_grid_dep_wait() blocks until that predecessor has
// completed and its writes are visible, nb_launch_dependents() lets the successor begin launching.  Both are
// no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void nb_grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void nb_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ unsigned long long nb_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Spin until *flag >= want (acquire, system scope), at most timeout_ns.  Returns false on timeout after
// recording (kind, peer, want) in *err_word.
__device__ __forceinline__ bool nb_wait_flag(const unsigned long long* flag, unsigned long long want,
                                             unsigned long long timeout_ns, unsigned long long* err_word,
                                             unsigned long long kind, int peer) {
    if (nb_ld_acquire_sys(flag) >= want) return true;
    const unsigned long long t0 = nb_globaltimer_ns();
    unsigned ns = 128;
    while (nb_ld_acquire_sys(flag) < want) {
        __nanosleep(ns);
        if (ns < 4096) ns <<= 1;
        if (nb_globaltimer_ns() - t0 > timeout_ns) {
            if (err_word) {
                *reinterpret_cast<volatile unsigned long long*>(err_word) =
                    (kind << 56) | ((unsigned long long)(peer & 0xff) << 48) | (want & 0xffffffffffffull);
                __threadfence_system();
            }
            return false;
        }
    }
    return true;
}

__device__ __forceinline__ float nb_rcp_f32(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // bare MUFU.RCP
    return y;
}
// 1/x for a normal positive double: MUFU.RCP64H seed (~2^-20) + one cubic step -> < 1.5 ulp.
__device__ __forceinline__ double nb_rcp_f64(double x) {
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    double e = fma(-x, y0, 1.0);
    double p = fma(e, e, e);
    return fma(y0, p, y0);
}
