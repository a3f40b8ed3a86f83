// nb200_api.cu -- the extern "C" layer of libnb200.so (see include/nb200.h for the contract and
// the reference file:line each entry point replaces).
//
// One nb200_ctx owns one Shard per device it drives: G shards in a single-process context
// (nb200_create), exactly one in a rank context (nb200_create_rank; one process per GPU under
// torchrun).  A shard owns a contiguous range of TARGET tiles, their FP64 master state, and a full
// double-buffered copy of the tile-planar SOURCES of all bodies.
//
// Per step and shard, default path (N >= ~32768, fused NVLink exchange attached or one shard):
//     [wait for the peers' step/epoch flags]  nb_wait_flags_kernel
//     close-pair pre-pass                      nb_grid_insert/query_kernel
//     pair-symmetric force pass                nb_force_sym_kernel   (launch_symmetric)
//     [reaction sums -> their owners' slots]   nb_sym_push_kernel    (peer stores + flag)
//     epilogue: sum, integrate, new rows to own AND peers' next buffers, step flag   nb_finish_kernel
// Small N / NCCL exchange / detached shards: the ordered pass nb_force_kernel with the integrator fused
// into its epilogue; with NCCL, pass A (own sources) overlaps the in-place ncclAllGather of the
// previous step's rows on the comm stream and pass B takes the other shards' sources.  NCCL is
// dlopen'ed on first use (no link-time dependency; a 1-GPU context never touches it).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/nb200.h"
#include "nb_aux.cuh"
#include "nb_force.cuh"
#include "nb_force_sym.cuh"
#include "nb_p2p.cuh"

#define NB200_VERSION_STR "nb200 0.1 (sm_100a)"

namespace {

thread_local std::string g_create_error;

// ------------------------------------------------------------------------------- NCCL (lazy)
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("NB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("cannot dlopen NCCL: ") + (dlerror() ? dlerror() : "?");
            return;
        }
#define NB_SYM(field, name)                                                    \
    *(void**)(&api.field) = dlsym(api.handle, name);                           \
    if (!api.field) { api.error = std::string("NCCL symbol missing: ") + name; return; }
        NB_SYM(GetUniqueId, "ncclGetUniqueId")
        NB_SYM(CommInitRank, "ncclCommInitRank")
        NB_SYM(CommInitAll, "ncclCommInitAll")
        NB_SYM(CommDestroy, "ncclCommDestroy")
        NB_SYM(AllGather, "ncclAllGather")
        NB_SYM(GroupStart, "ncclGroupStart")
        NB_SYM(GroupEnd, "ncclGroupEnd")
        NB_SYM(GetErrorString, "ncclGetErrorString")
#undef NB_SYM
    });
    return &api;
}

// ------------------------------------------------------------------------------- kernel table
struct Variant {
    int ti, js, block;
    int itile() const { return block / js * ti; }
};
// index = "variant" option.  Large i-tiles first (the auto-planner prefers the largest that still
// yields enough work units).
const Variant kVariants[] = {
    {4, 1, 256},   // 0: 1024 targets per i-tile
    {2, 1, 256},   // 1:  512
    {4, 4, 256},   // 2:  256
    {2, 4, 128},   // 3:   64
    {2, 8, 128},   // 4:   32
    // low-ILP, high-occupancy shapes: few chains per thread keep the three accumulate FFMA2 of a
    // chain adjacent, so ptxas marks their shared multiplier .reuse (one operand fetch saved each)
    {2, 1, 256},   // 5:  512, inner loop not unrolled
    {1, 1, 256},   // 6:  256
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr int kNumAutoVariants = 5;     // the planner picks among 0..4; the rest are opt-in ("variant" option)

typedef void (*ForceKernel)(const NbForceParams);

template <int D, bool F64, bool FLAGS> ForceKernel kernel_for(int v) {
    switch (v) {
        case 0: return nb_force_kernel<D, F64, 4, 1, 256, FLAGS>;
        case 1: return nb_force_kernel<D, F64, 2, 1, 256, FLAGS>;
        case 2: return nb_force_kernel<D, F64, 4, 4, 256, FLAGS>;
        case 3: return nb_force_kernel<D, F64, 2, 4, 128, FLAGS>;
        case 5: return nb_force_kernel<D, F64, 2, 1, 256, FLAGS, 1>;
        case 6: return nb_force_kernel<D, F64, 1, 1, 256, FLAGS, 2>;
        default: return nb_force_kernel<D, F64, 2, 8, 128, FLAGS>;
    }
}
// flags = the close-pair pre-pass ran and P.suspect is valid (NB_PLAIN / NB_EXACT per warp and tile)
ForceKernel pick_kernel(int dim, bool f64, int v, bool flags = false) {
    if (flags) {
        if (dim == 3) return f64 ? kernel_for<3, true, true>(v) : kernel_for<3, false, true>(v);
        return f64 ? kernel_for<2, true, true>(v) : kernel_for<2, false, true>(v);
    }
    if (dim == 3) return f64 ? kernel_for<3, true, false>(v) : kernel_for<3, false, false>(v);
    return f64 ? kernel_for<2, true, false>(v) : kernel_for<2, false, false>(v);
}

size_t smem_bytes(int dim, bool f64) {
    const size_t tile = (size_t)NB_TILE * (dim + 1) * (f64 ? 8 : 4);
    return NB_STAGES * tile + 2 * NB_STAGES * sizeof(uint64_t) + 16;
}

// ------------------------------------------------------------------------------- state
struct Shard {
    int device = 0;
    int rank = 0;
    long long tile_lo = 0, tile_hi = 0;   // owned source/target tiles
    long long tgt_base = 0;               // tile_lo * NB_TILE
    long long n_local = 0;                // real bodies owned
    int tpad = 0;                         // padded targets (multiple of the largest i-tile)
    cudaStream_t compute = nullptr, comm = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_pass_done = nullptr;
    cudaEvent_t ev_gather[2] = {nullptr, nullptr};
    void* src[2] = {nullptr, nullptr};
    double *acc = nullptr, *pos = nullptr, *vel = nullptr, *mass = nullptr, *forces = nullptr;
    double* aos_dev = nullptr;            // staging image of the AoS bodies (upload/download)
    size_t aos_bytes = 0;
    unsigned long long* bounds = nullptr; // [3] bit patterns of max |coordinate|, max |mass|, min |mass| of the image
    double* energy = nullptr;             // [2]
    double* cmp = nullptr;                // [2][n_local*D] staging of host force arrays for nb200_accuracy_pct
    bool forces_valid = false;            // s.forces holds the result of the last nb200_forces call
    unsigned *tile_done = nullptr, *sched = nullptr;
    ncclComm_t comm_nccl = nullptr;
    int sms = 0;
    // close-pair pre-pass
    unsigned long long* grid_keys = nullptr;
    unsigned* grid_counts = nullptr;
    unsigned grid_cap = 0;
    unsigned char* suspect = nullptr;                 // [tpad]
    // pair-symmetric pass (nb_force_sym.cuh): work list, global-index accumulators, counters
    NbSymRow* sym_rows = nullptr;
    int* sym_prefix = nullptr;
    int sym_rows_cap = 0;
    std::vector<NbSymRow> sym_rows_host;
    std::vector<int> sym_prefix_host;
    int sym_key_seg = 0, sym_key_world = 0, sym_key_tpi = 0;
    float* src_lo[2] = {nullptr, nullptr};   // 48-bit positions: lo parts of the scaled coordinates, tile-planar [D][256]
    const void* occ_fn = nullptr;        // pair-symmetric kernel of the last occupancy query and its answer
    int occ_blocks = 0;
    double* gacc = nullptr;                           // [3][nalloc]
    unsigned* sym_done = nullptr;                     // [2] push / finish CTA counters
    // fused NVLink exchange (peer stores from the epilogue + flag handshake)
    unsigned long long* flags = nullptr;              // [2*kMaxWorldP2P]: step flags, then epoch flags, by writer rank
    int n_peers = 0;
    int peer_rank[NB_MAX_PEERS] = {0};
    void* peer_src[2][NB_MAX_PEERS] = {{nullptr}};    // peers' source buffers, mapped into this device
    unsigned long long* peer_flags[NB_MAX_PEERS] = {nullptr};
    bool ipc_opened = false;                          // peer pointers came from cudaIpcOpenMemHandle
    double* det_slots = nullptr;                      // "deterministic" option: [segments][3][tpad] unit partial sums
    size_t det_slots_doubles = 0;
    unsigned long long* err_host = nullptr;           // page-locked, device-mapped: what a timed-out flag wait was waiting for
    unsigned long long* err_dev = nullptr;            // its device address
};

constexpr int kMaxWorldP2P = NB_MAX_PEERS + 1;
constexpr int kSymMaxSlots = kMaxWorldP2P / 2;      // senders of reaction sums per rank: floor(world / 2)
// One allocation per shard (one IPC handle): 4 * kMaxWorldP2P flag words (step, upload epoch, reaction-sum pass, upload
// ready; one per writer rank each), the per-rank bounds table of a shard-local upload (3 words per rank: max |x|, max |m|,
// min |m|), then the
// receive slots of the pair-symmetric pass.
constexpr size_t kFlagsBytes = 512;
constexpr int kBoundsTableWord = 4 * (NB_MAX_PEERS + 1);          // 3 words per rank: 32 + 24 <= 64 words
// The 256-target i-tile shape is opt-in only ("sym_itile" option): measured, it never beats the 1024
// shape nor, below N ~ 32768, the ordered pass (N=16384: 1414 vs 1454 vs 1994 G inter/s).
constexpr size_t kSymSmallN = 0;
// Reaction-sum reduction of the pair-symmetric kernel: 0 shared-memory transpose, 1 register rotation through the
// warp, 2 (FP32) rotation with decoupled hand-over.  Measured at N = 2^20, FP32 3D: 3323 / 3811 / 3757 G inter/s
// with 4 x 256 threads, 3864 with 8 x 128 and rotation (profiles/r02_sym_variants.jsonl).
constexpr int kSymAlgoDefault = 2;     // FP64 has no flavour 2: it takes 1
constexpr int kSymTiF32 = 8;
constexpr int kSymTiF64 = 8;     // 8 x 128 on two CTAs per SM: 1707 vs 1584 G inter/s (4 x 256) at N = 2^20, 3D

}  // namespace

struct nb200_ctx {
    int dim = 3;
    size_t n = 0;
    bool f64 = true;
    int world = 1;                // shards over all processes
    bool rank_mode = false;
    bool detached = false;        // rank context without a communicator (single-GPU test hook)
    bool p2p_ready = false;       // peer pointers + flags of every peer are mapped
    long long ntiles = 0;         // source tiles incl. padding: world * tiles_per_shard
    long long tiles_per_shard = 0;
    long long nalloc = 0;         // bodies allocated in each source buffer
    std::vector<Shard> shards;    // the shards THIS process drives
    int cur = 0;                  // current source buffer
    bool uploaded = false;
    double pos_scale = 1.0, mass_scale = 1.0;
    double xmax = 1.0;            // max |coordinate| at upload (cell size of the close-pair grid when cutoff = 0)
    int exchange = 0;             // 0 = NCCL all-gather, 1 = peer stores fused into the epilogue
    unsigned long long step_index = 0;   // steps issued since creation (the published flag value)
    unsigned long long epoch_base = 0;   // step_index at the last upload
    unsigned long long epoch = 0;        // uploads so far (published on the epoch flags)
    std::vector<unsigned long long> acc_seq_issued;   // per driven shard: pair-symmetric passes with a reaction exchange so far
    bool pristine = false;        // no step since the last upload: the AoS staging image is still current
    bool image_full = true;       // the staging image holds all n bodies (false: only the rows each shard owns)
    bool equal_mass = false;      // every body has the same (positive) mass: the pair kernels run without the masses
    double common_mass = 0.0;
    bool dead = false;            // a peer handshake timed out: the ranks' step counters may have diverged
    // options
    int opt_variant = -1, opt_seg_tiles = 0, opt_grid_mult = 0, opt_overlap = -1, opt_trace = 0, opt_detect = -1, opt_symmetric = -1, opt_sym_ti = 0, opt_sym_itile = 0, opt_sym_algo = -1, opt_sym_block = 0, opt_seg_sub = 0, opt_shard_upload = -1, opt_pdl = -1, opt_eqm = -1, opt_deterministic = 0;
    bool opt_pos48 = false;          // FP32 mode with 48-bit positions (hi + lo float pairs): option "fp32_positions"
    long opt_spin_timeout_ms = 30000;   // bound of every device-side wait on a peer's flag
    std::vector<std::pair<std::string, cudaEvent_t>> trace;   // shard-0 timeline of the last step call (opt_trace)
    // bookkeeping
    long long launches = 0;
    double last_ms = 0.0;
    std::string error, plan;
    size_t aos_stride = 0;
};

namespace {

int fail(nb200_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->error = buf;
    else g_create_error = buf;
    return code;
}

#define CK(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(ctx, NB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                   \
    } while (0)

#define CKN(call)                                                                              \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess)                                                                 \
            return fail(ctx, NB200_ENCCL, "%s failed: %s (%s:%d)", #call,                       \
                        nccl_api()->GetErrorString(r_), __FILE__, __LINE__);                   \
    } while (0)

constexpr int kMaxItile = 1024;

typedef void (*SymKernel)(const NbSymParams);
// pair-symmetric kernels (nb_force_sym.cuh)
// One shape of the pair-symmetric kernel: targets per thread x threads (i-tile = ti * block targets).
struct SymShape { int ti, block; };
template <int D, bool F64, int ALGO> SymKernel sym_kernel_of(SymShape sh) {
    if (sh.block == 64) return nb_force_sym_kernel<D, F64, 4, 64, 0>;            // small-N experiment, transpose only
    if constexpr (F64) {
        if (sh.ti == 2) return sh.block == 128 ? nb_force_sym_kernel<D, true, 2, 128, ALGO> : nb_force_sym_kernel<D, true, 2, 256, ALGO>;
        if constexpr (ALGO == 1) {
            if (sh.ti == 8) return nb_force_sym_kernel<D, true, 8, 128, 1>;
            if (sh.block == 128) return nb_force_sym_kernel<D, true, 4, 128, 1>;
        }
        return nb_force_sym_kernel<D, true, 4, 256, ALGO>;
    } else {
        if (sh.ti == 8) return nb_force_sym_kernel<D, false, 8, 128, ALGO>;
        return sh.block == 128 ? nb_force_sym_kernel<D, false, 4, 128, ALGO> : nb_force_sym_kernel<D, false, 4, 256, ALGO>;
    }
}
template <int D, bool F64> SymKernel sym_kernel_of(SymShape sh, int algo) {
    if (algo == 1) return sym_kernel_of<D, F64, 1>(sh);
    if constexpr (!F64) { if (algo == 2) return sym_kernel_of<D, false, 2>(sh); }
    return sym_kernel_of<D, F64, 0>(sh);
}
// The equal-mass flavour exists for the default shapes and reduction only (FP32: 8 x 128 and 4 x 128 with the decoupled
// rotation; FP64: 4 x 256 with the rotation); every other combination runs the general kernels, which are correct for
// equal masses too.
bool sym_has_eqm(bool f64, SymShape sh, int algo) {
    if (f64) return algo == 1 && ((sh.ti == 4 && sh.block == 256) || (sh.ti == 8 && sh.block == 128));
    return algo == 2 && sh.block == 128 && (sh.ti == 8 || sh.ti == 4);
}
template <int D> SymKernel sym_kernel_eqm(bool f64, SymShape sh) {
    if (f64) return sh.ti == 8 ? nb_force_sym_kernel<D, true, 8, 128, 1, true> : nb_force_sym_kernel<D, true, 4, 256, 1, true>;
    return sh.ti == 8 ? nb_force_sym_kernel<D, false, 8, 128, 2, true> : nb_force_sym_kernel<D, false, 4, 128, 2, true>;
}
SymKernel pick_sym_kernel(int dim, bool f64, SymShape sh, int algo, bool eqm = false, bool hl = false) {
    if (hl) return dim == 3 ? nb_force_sym_kernel<3, false, 8, 128, 2, false, true> : nb_force_sym_kernel<2, false, 8, 128, 2, false, true>;
    if (eqm && sym_has_eqm(f64, sh, algo)) return dim == 3 ? sym_kernel_eqm<3>(f64, sh) : sym_kernel_eqm<2>(f64, sh);
    if (dim == 3) return f64 ? sym_kernel_of<3, true>(sh, algo) : sym_kernel_of<3, false>(sh, algo);
    return f64 ? sym_kernel_of<2, true>(sh, algo) : sym_kernel_of<2, false>(sh, algo);
}
// every shape this precision can be launched with (for the one-time shared-memory opt-in)
const SymShape kSymShapesF32[] = {{4, 256}, {8, 128}, {4, 128}, {4, 64}};
const SymShape kSymShapesF64[] = {{4, 256}, {2, 256}, {2, 128}, {8, 128}, {4, 128}, {4, 64}};

// CUDA loads a kernel's code lazily at its first launch; the reference times ONE call per process
// (safely_execute, utils.h:87-104), so every kernel a call may launch is loaded when the context is created
// (querying a function's attributes loads it).  The force kernels are covered by their cudaFuncSetAttribute calls.
template <int D, typename real> int preload_aux_kernels_t(nb200_ctx* ctx) {
    const void* fns[] = {
        (const void*)nb_pack_kernel<D, real>,        (const void*)nb_bounds_kernel<D>,        (const void*)nb_unpack_kernel<D>,
        (const void*)nb_grid_insert_kernel<D, real>, (const void*)nb_grid_query_kernel<D, real>,
        (const void*)nb_finish_kernel<D, real>,      (const void*)nb_sym_push_kernel<D>,      (const void*)nb_energy_kernel<D, real>,
        (const void*)nb_accuracy_kernel<D>,          (const void*)nb_generate_kernel<D>,      (const void*)nb_compare_kernel<D>,
    };
    cudaFuncAttributes at;
    for (const void* f : fns) CK(cudaFuncGetAttributes(&at, f));
    return NB200_OK;
}
int preload_aux_kernels(nb200_ctx* ctx) {
    if (ctx->dim == 3) return ctx->f64 ? preload_aux_kernels_t<3, double>(ctx) : preload_aux_kernels_t<3, float>(ctx);
    return ctx->f64 ? preload_aux_kernels_t<2, double>(ctx) : preload_aux_kernels_t<2, float>(ctx);
}

int alloc_shard(nb200_ctx* ctx, Shard& s) {
    const int D = ctx->dim;
    const size_t rs = ctx->f64 ? 8 : 4;
    CK(cudaSetDevice(s.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, s.device));
    if (prop.major < 10)
        return fail(ctx, NB200_ECUDA, "device %d is sm_%d%d; libnb200 is built for sm_100a only", s.device,
                    prop.major, prop.minor);
    s.sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&s.compute, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s.comm, cudaStreamNonBlocking));
    CK(cudaEventCreate(&s.ev_start));
    CK(cudaEventCreate(&s.ev_stop));
    CK(cudaEventCreateWithFlags(&s.ev_pass_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_gather[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_gather[1], cudaEventDisableTiming));
    const size_t src_bytes = (size_t)ctx->nalloc * (D + 1) * rs;
    for (int b = 0; b < 2; ++b) CK(cudaMalloc(&s.src[b], src_bytes));
    const size_t tp = (size_t)s.tpad;
    CK(cudaMalloc(&s.acc, 3 * tp * sizeof(double)));
    CK(cudaMemset(s.acc, 0, 3 * tp * sizeof(double)));
    CK(cudaMalloc(&s.pos, (size_t)D * tp * sizeof(double)));
    CK(cudaMalloc(&s.vel, (size_t)D * tp * sizeof(double)));
    CK(cudaMalloc(&s.mass, tp * sizeof(double)));
    CK(cudaMalloc(&s.forces, std::max<size_t>(1, (size_t)s.n_local * D) * sizeof(double)));
    // staging image of the AoS bodies, sized for the packed Body<D> record (40 / 56 bytes) so that no allocation is left
    // for the first upload -- the reference times ONE call per process; a larger stride reallocates in nb200_upload_aos
    s.aos_bytes = std::max<size_t>(1, ctx->n) * (size_t)(2 * D + 1) * sizeof(double);
    CK(cudaMalloc(&s.aos_dev, s.aos_bytes));
    CK(cudaMalloc(&s.energy, 2 * sizeof(double)));
    CK(cudaMalloc(&s.bounds, 3 * sizeof(unsigned long long)));
    CK(cudaMalloc(&s.tile_done, (tp / 32 + 1) * sizeof(unsigned)));
    CK(cudaMemset(s.tile_done, 0, (tp / 32 + 1) * sizeof(unsigned)));
    {
        unsigned cap = 1024;
        while ((long long)cap < 2 * ctx->ntiles * NB_TILE) cap <<= 1;
        s.grid_cap = cap;
        // keys and counts in one allocation: one memset per step clears both
        CK(cudaMalloc(&s.grid_keys, (size_t)cap * (sizeof(unsigned long long) + sizeof(unsigned))));
        s.grid_counts = reinterpret_cast<unsigned*>(s.grid_keys + cap);
        CK(cudaMalloc(&s.suspect, tp));
        CK(cudaMemset(s.suspect, 1, tp));
        {
            // rows: per own i-tile one ordered + one triangular row, plus one row per cross-shard block
            s.sym_rows_cap = (int)(tp / NB_TILE + 1) * (3 + kSymMaxSlots);     // i-tiles may be as small as one source tile
            CK(cudaMalloc(&s.sym_rows, (size_t)s.sym_rows_cap * sizeof(NbSymRow)));
            CK(cudaMalloc(&s.sym_prefix, (size_t)(s.sym_rows_cap + 1) * sizeof(int)));
            CK(cudaMalloc(&s.gacc, 3 * (size_t)ctx->nalloc * sizeof(double)));
            CK(cudaMemset(s.gacc, 0, 3 * (size_t)ctx->nalloc * sizeof(double)));
            CK(cudaMalloc(&s.sym_done, 2 * sizeof(unsigned)));
            CK(cudaMemset(s.sym_done, 0, 2 * sizeof(unsigned)));
        }
    }
    {
        // flag words (step, epoch, reaction-sum pass: one per writer rank each) and, behind them in the
        // SAME allocation (one IPC handle), the receive slots of the pair-symmetric pass
        const size_t slots = (size_t)(ctx->world / 2) * 3 * (size_t)ctx->tiles_per_shard * NB_TILE * sizeof(double);   // floor(G/2) senders
        CK(cudaMalloc(&s.flags, kFlagsBytes + slots));
        CK(cudaMemset(s.flags, 0, kFlagsBytes));
    }
    CK(cudaMalloc(&s.sched, 2 * sizeof(unsigned)));
    CK(cudaMemset(s.sched, 0, 2 * sizeof(unsigned)));
    CK(cudaHostAlloc(reinterpret_cast<void**>(&s.err_host), sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable));
    *s.err_host = 0ull;
    CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&s.err_dev), s.err_host, 0));
    // opt in to the dynamic shared memory of every variant once
    for (int v = 0; v < kNumVariants; ++v)
        for (int fl = 0; fl < 2; ++fl)
            CK(cudaFuncSetAttribute((const void*)pick_kernel(D, ctx->f64, v, fl != 0),
                                    cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_bytes(D, ctx->f64)));
    const SymShape* shapes = ctx->f64 ? kSymShapesF64 : kSymShapesF32;
    const size_t n_shapes = ctx->f64 ? sizeof kSymShapesF64 / sizeof(SymShape) : sizeof kSymShapesF32 / sizeof(SymShape);
    for (size_t k = 0; k < n_shapes; ++k) {
        const SymShape sh = shapes[k];
        const bool rot_only = ctx->f64 && sh.block == 128 && sh.ti >= 4;
        for (int algo = rot_only ? 1 : 0; algo <= (sh.block == 64 ? 0 : ctx->f64 ? 1 : 2); ++algo)
            CK(cudaFuncSetAttribute((const void*)pick_sym_kernel(D, ctx->f64, sh, algo), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)nb_sym_smem_bytes(D, sh.block, ctx->f64, sh.ti, algo)));
    }
    for (size_t k = 0; k < n_shapes; ++k) {
        const SymShape sh = shapes[k];
        for (int algo = 1; algo <= 2; ++algo)
            if (sym_has_eqm(ctx->f64, sh, algo))
                CK(cudaFuncSetAttribute((const void*)pick_sym_kernel(D, ctx->f64, sh, algo, true), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)nb_sym_smem_bytes(D, sh.block, ctx->f64, sh.ti, algo)));
    }
    if (!ctx->f64)
        CK(cudaFuncSetAttribute((const void*)pick_sym_kernel(D, false, SymShape{8, 128}, 2, false, true), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)nb_sym_smem_bytes(D, 128, false, 8, 2, true)));
    return preload_aux_kernels(ctx);
}

void free_shard(Shard& s) {
    cudaSetDevice(s.device);
    if (s.compute) cudaStreamSynchronize(s.compute);
    if (s.comm) cudaStreamSynchronize(s.comm);
    if (s.comm_nccl && nccl_api()->CommDestroy) nccl_api()->CommDestroy(s.comm_nccl);
    if (s.ipc_opened) {
        for (int p = 0; p < s.n_peers; ++p) {
            cudaIpcCloseMemHandle(s.peer_src[0][p]);
            cudaIpcCloseMemHandle(s.peer_src[1][p]);
            cudaIpcCloseMemHandle(s.peer_flags[p]);
        }
    }
    cudaFree(s.flags);
    if (s.err_host) cudaFreeHost(s.err_host);
    cudaFree(s.det_slots);
    cudaFree(s.grid_keys); cudaFree(s.suspect); cudaFree(s.sym_rows); cudaFree(s.sym_prefix); cudaFree(s.gacc); cudaFree(s.sym_done);
    for (int b = 0; b < 2; ++b) cudaFree(s.src[b]);
    for (int b = 0; b < 2; ++b) cudaFree(s.src_lo[b]);
    cudaFree(s.acc); cudaFree(s.pos); cudaFree(s.vel); cudaFree(s.mass); cudaFree(s.forces);
    cudaFree(s.aos_dev); cudaFree(s.energy); cudaFree(s.bounds); cudaFree(s.cmp); cudaFree(s.tile_done); cudaFree(s.sched);
    if (s.ev_start) cudaEventDestroy(s.ev_start);
    if (s.ev_stop) cudaEventDestroy(s.ev_stop);
    if (s.ev_pass_done) cudaEventDestroy(s.ev_pass_done);
    for (int b = 0; b < 2; ++b) if (s.ev_gather[b]) cudaEventDestroy(s.ev_gather[b]);
    if (s.compute) cudaStreamDestroy(s.compute);
    if (s.comm) cudaStreamDestroy(s.comm);
}

// geometry shared by both create paths
int layout(nb200_ctx* ctx, int dim, size_t n, int precision, int world) {
    if (dim != 2 && dim != 3) return fail(nullptr, NB200_EINVAL, "dim must be 2 or 3 (main.cpp:889-892), got %d", dim);
    if (precision != NB200_FP64 && precision != NB200_FP32)
        return fail(nullptr, NB200_EINVAL, "precision must be 64 or 32, got %d", precision);
    if (world < 1 || world > 64) return fail(nullptr, NB200_EINVAL, "bad shard count %d", world);
    if (n > (size_t)1 << 30) return fail(nullptr, NB200_EINVAL, "n too large");
    ctx->dim = dim;
    ctx->n = n;
    ctx->f64 = precision == NB200_FP64;
    ctx->world = world;
    const long long tiles = std::max<long long>(1, ((long long)n + NB_TILE - 1) / NB_TILE);
    ctx->tiles_per_shard = (tiles + world - 1) / world;
    ctx->ntiles = ctx->tiles_per_shard * world;
    // slack so that a padded i-tile of the last shard never reads past the buffer
    ctx->nalloc = ctx->ntiles * NB_TILE + kMaxItile;
    return NB200_OK;
}

void place_shard(const nb200_ctx* ctx, Shard& s, int rank, int device) {
    s.rank = rank;
    s.device = device;
    s.tile_lo = rank * ctx->tiles_per_shard;
    s.tile_hi = s.tile_lo + ctx->tiles_per_shard;
    s.tgt_base = s.tile_lo * NB_TILE;
    const long long hi = std::min<long long>((long long)ctx->n, s.tile_hi * NB_TILE);
    s.n_local = std::max<long long>(0, hi - s.tgt_base);
    const long long span = ctx->tiles_per_shard * NB_TILE;
    s.tpad = (int)((span + kMaxItile - 1) / kMaxItile * kMaxItile);
}

// ------------------------------------------------------------------------------- launch plan
struct Plan {
    int variant;
    int seg_tiles;
    int grid;
    int n_itiles;
    bool flags;       // close-pair pre-pass + NB_PLAIN/NB_EXACT kernels instead of the tracked pass
};

// The close-pair pre-pass (two memsets + two small launches per step, ~70 us) lets the force kernels drop all per-pair
// cut-off work.  Against the exact-cut-off flavour of the pair-symmetric pass (~10 % slower chains in 3D, ~15 % in 2D) it
// pays from N ~ 50000: measured at N = 65536, FP32: 1.275 vs 1.296 ms/step in 3D, 0.944 vs 1.002 in 2D; at 32768 (3D) it
// loses, 0.432 vs 0.351 (profiles/r02/small_n.md).
bool use_detect(const nb200_ctx* ctx) {
    if (ctx->opt_detect >= 0) return ctx->opt_detect != 0;
    return ctx->n >= 49152u;
}

int make_plan(nb200_ctx* ctx, const Shard& s, Plan* out) {
    const int D = ctx->dim;
    const long long NT = ctx->ntiles;
    int occ = 1;
    const bool flags = use_detect(ctx);
    out->flags = flags;
    auto resident = [&](int v, int* grid) -> int {
        int nb = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &nb, (const void*)pick_kernel(D, ctx->f64, v, flags), kVariants[v].block, smem_bytes(D, ctx->f64));
        if (e != cudaSuccess) return -1;
        *grid = std::max(1, nb) * s.sms;
        return nb;
    };
    const long long span = ctx->tiles_per_shard * NB_TILE;   // targets incl. tile padding
    int v = ctx->opt_variant;
    int grid = 0;
    if (ctx->opt_deterministic) {
        // one shape and one segmentation for every shard count: each target's sum is the same expression on 1 and on N GPUs
        v = (v >= 0 && v < kNumVariants) ? v : 0;
    } else if (v < 0 || v >= kNumVariants) {
        // largest i-tile that still gives every resident CTA >= 8 units at >= 8 tiles per unit,
        // else the smallest i-tile
        v = kNumAutoVariants - 1;
        for (int c = 0; c < kNumAutoVariants; ++c) {
            int g = 0;
            if (resident(c, &g) < 0) continue;
            const long long nit = (span + kVariants[c].itile() - 1) / kVariants[c].itile();
            const long long max_units = nit * std::max<long long>(1, NT / 8);
            if (max_units >= 8LL * g) { v = c; break; }
        }
    }
    occ = resident(v, &grid);
    if (occ < 0) return fail(ctx, NB200_ECUDA, "occupancy query failed for variant %d", v);
    const int itile = kVariants[v].itile();
    const int nit = (int)((span + itile - 1) / itile);
    int seg = ctx->opt_seg_tiles;
    if (seg <= 0 && ctx->opt_deterministic) seg = 16;
    if (seg <= 0) {
        // aim for ~64 units per resident CTA, at least 8 tiles per unit when the problem allows
        const long long want_units = 64LL * grid;
        long long nseg = std::max<long long>(1, (want_units + nit - 1) / nit);
        nseg = std::min<long long>(nseg, std::max<long long>(1, NT / 8));
        if ((long long)nit * nseg < 2LL * grid) nseg = std::min<long long>(NT, std::max<long long>(nseg, (2LL * grid + nit - 1) / nit));
        seg = (int)((NT + nseg - 1) / nseg);
    }
    seg = (int)std::max<long long>(1, std::min<long long>(seg, NT));
    if (ctx->opt_grid_mult > 0) grid = std::max(1, grid * ctx->opt_grid_mult / 16);
    out->variant = v;
    out->seg_tiles = seg;
    out->grid = grid;
    out->n_itiles = nit;
    return NB200_OK;
}

struct Ranges {
    int b[3], e[3];
    Ranges(int b0, int e0, int b1 = 0, int e1 = 0, int b2 = 0, int e2 = 0) : b{b0, b1, b2}, e{e0, e1, e2} {}
};

int nsegs(int b, int e, int seg) { return e > b ? (e - b + seg - 1) / seg : 0; }

// Launch with programmatic dependent launch allowed: the kernel may start (and run its prologue) while its predecessor in
// the stream drains; it waits for the predecessor itself (nb_grid_dep_wait).  A step of a small problem is two short
// launches back to back: the launch latency between them is the part of the step this hides.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// launch one pass of the force kernel on shard s over the given source-tile ranges
// close-pair pre-pass on the shard's compute stream: hash-grid insert of every source, then the
// suspect flag of every own (padded) target.  Reads src[cur], so it runs after the peer handshake.
int launch_detect(nb200_ctx* ctx, Shard& s, double cutoff, int cur) {
    const int D = ctx->dim;
    // every source takes part in the grid, the zero-mass padding included (it sits on a real body's position); parked
    // padding (equal-mass systems) is out of range of everything and stays out of the grid
    const long long nbodies = ctx->equal_mass ? (long long)ctx->n : ctx->ntiles * NB_TILE;
    const double cs = cutoff * ctx->pos_scale * ctx->pos_scale;
    NbGrid g;
    g.keys = s.grid_keys;
    g.counts = s.grid_counts;
    g.mask = s.grid_cap - 1;
    // cell edge >= sqrt(cutoff) * 1.001; with no cut-off only exact duplicates matter: any tiny cell does,
    // but keep cell indices far inside the int64 range
    // 48-bit positions: the grid is built on the hi parts, whose pair differences are off by up to 2^-24 per coordinate
    // (|x'| <= 1 at upload): cells twice that much wider keep "no other body in the 3^D cells around" a proof that no TRUE distance is
    // under the cut-off
    const double h = (cs > 0.0 ? sqrt(cs) * 1.001 : ldexp(ctx->xmax * ctx->pos_scale, -40)) +
                     (ctx->opt_pos48 ? ldexp(std::max(1.0, ctx->xmax * ctx->pos_scale), -23) : 0.0);   // 2x: bodies may leave the uploaded range
    g.inv_h = 1.0 / h;
    // one memset (keys = empty, counts = -1), then insert -> query -> force pass chained by programmatic dependent launch
    CK(cudaMemsetAsync(s.grid_keys, 0xFF, (size_t)s.grid_cap * (sizeof(unsigned long long) + sizeof(unsigned)), s.compute));
    const int threads = 256;
    const int bi = (int)((nbodies + threads - 1) / threads), bq = (s.tpad + threads - 1) / threads;
    const bool pdl = ctx->opt_pdl != 0;
#define NB_GRID(DD, RR)                                                                                                      \
    CK(launch_pdl(nb_grid_insert_kernel<DD, RR>, bi, threads, 0, s.compute, false, (const RR*)s.src[cur], nbodies, g));      \
    CK(launch_pdl(nb_grid_query_kernel<DD, RR>, bq, threads, 0, s.compute, pdl, (const RR*)s.src[cur], s.tgt_base, s.tpad,   \
                  nbodies, g, s.suspect))
    if (D == 3) { if (ctx->f64) { NB_GRID(3, double); } else { NB_GRID(3, float); } }
    else        { if (ctx->f64) { NB_GRID(2, double); } else { NB_GRID(2, float); } }
#undef NB_GRID
    ctx->launches += 2;
    return NB200_OK;
}

struct Handshake {
    bool exchange = false;       // fused peer-store exchange: epilogue rows also go to the peers' next buffers
    bool lazy = false;           // range 0 = own rows only: handshake when a CTA first reaches a remote unit
    unsigned long long wait_step = 0, wait_epoch = 0, signal_step = 0;
};

struct PeerRanks {
    int r[NB_MAX_PEERS];
    explicit PeerRanks(const Shard& s) { for (int p = 0; p < NB_MAX_PEERS; ++p) r[p] = s.peer_rank[p]; }
};

// one warp: lane p spins until peer p has published step >= want_step and epoch >= want_epoch
__global__ void nb_wait_flags_kernel(const unsigned long long* flags, int stride, int n_peers, PeerRanks pr,
                                     unsigned long long want_step, unsigned long long want_epoch,
                                     unsigned long long timeout_ns, unsigned long long* err_word) {
    const int p = threadIdx.x;
    if (p < n_peers) {
        nb_wait_flag(flags + pr.r[p], want_step, timeout_ns, err_word, NB_WAIT_STEP, pr.r[p]) &&
            nb_wait_flag(flags + stride + pr.r[p], want_epoch, timeout_ns, err_word, NB_WAIT_EPOCH, pr.r[p]);
    }
}

// shard-local upload, step 1: this shard's bounds (max |x|, max |m| of its own rows) go into slot [rank] of every
// shard's bounds table, then the "ready" flag tells the peers that (a) the bounds are there and (b) this shard has
// finished every kernel of the previous upload epoch (stream order), so its source buffers may be overwritten
__global__ void nb_publish_ready_kernel(NbForceParams P, const unsigned long long* my_bounds, int table_word, int ready_word,
                                        unsigned long long epoch) {
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) {                 // max |x|, max |m|, min |m| of the own rows
            const unsigned long long v = my_bounds[k];
            P.my_flags[table_word + 3 * P.my_rank + k] = v;
            for (int p = 0; p < P.n_peers; ++p) P.peer_flags[p][table_word + 3 * P.my_rank + k] = v;
        }
        __threadfence_system();
        for (int p = 0; p < P.n_peers; ++p) nb_st_release_sys(P.peer_flags[p] + ready_word + P.my_rank, epoch);
    }
}
// step 2: wait until every peer is ready for upload epoch `epoch`
__global__ void nb_wait_ready_kernel(const unsigned long long* flags, int ready_word, int n_peers, PeerRanks pr,
                                     unsigned long long epoch, unsigned long long timeout_ns, unsigned long long* err_word) {
    const int p = threadIdx.x;
    if (p < n_peers) nb_wait_flag(flags + ready_word + pr.r[p], epoch, timeout_ns, err_word, NB_WAIT_READY, pr.r[p]);
}

// publish this shard's epoch (upload count) on every peer's epoch flag
__global__ void nb_publish_epoch_kernel(NbForceParams P, int stride, unsigned long long epoch) {
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (int p = 0; p < P.n_peers; ++p) nb_st_release_sys(P.peer_flags[p] + stride + P.my_rank, epoch);
    }
}

// everything of NbForceParams that does not depend on the unit decomposition of one launch
NbForceParams base_params(const nb200_ctx* ctx, const Shard& s, int mode, double G, double cutoff, double dt, int cur,
                          const Handshake& hs) {
    NbForceParams P;
    memset(&P, 0, sizeof P);
    P.src = s.src[cur];
    P.src_next = s.src[cur ^ 1];
    P.acc = s.acc;
    P.tile_done = s.tile_done;
    P.sched = s.sched;
    P.pos = s.pos;
    P.vel = s.vel;
    P.mass = s.mass;
    P.forces = s.forces;
    P.tgt_base = s.tgt_base;
    P.n_local = s.n_local;
    P.tpad = s.tpad;
    P.mode = mode;
    P.G = G;
    // FP32 sources are scaled by powers of two: x' = x ps, m' = m ms  =>  r'^2 = r^2 ps^2,
    // sum' = sum * ms / ps^3
    P.cutoff = cutoff * ctx->pos_scale * ctx->pos_scale;
    P.dt = dt;
    P.acc_scale = (ctx->pos_scale * ctx->pos_scale * ctx->pos_scale) / ctx->mass_scale;
    P.pos_scale = ctx->pos_scale;
    P.my_flags = s.flags;
    P.my_rank = s.rank;
    P.wait_step = hs.wait_step;
    P.wait_epoch = hs.wait_epoch;
    P.signal_step = hs.signal_step;
    P.flag_stride = kMaxWorldP2P;
    P.err_word = s.err_dev;
    P.spin_timeout_ns = (unsigned long long)std::max(1L, ctx->opt_spin_timeout_ms) * 1000000ull;
    if (hs.exchange) {
        P.n_peers = s.n_peers;
        for (int p = 0; p < s.n_peers; ++p) {
            P.peer_next[p] = s.peer_src[cur ^ 1][p];
            P.peer_flags[p] = s.peer_flags[p];
            P.peer_rank[p] = s.peer_rank[p];
        }
    }
    return P;
}

int launch_pass(nb200_ctx* ctx, Shard& s, const Plan& pl, const Ranges& rg, unsigned units_per_itile,
                int mode, double G, double cutoff, double dt, int cur, const Handshake& hs = Handshake()) {
    NbForceParams P = base_params(ctx, s, mode, G, cutoff, dt, cur, hs);
    P.suspect = pl.flags ? s.suspect : nullptr;
    P.n_itiles = pl.n_itiles;
    P.seg_tiles = pl.seg_tiles;
    int nseg_total = 0;
    for (int r = 0; r < 3; ++r) {
        P.rng_begin[r] = rg.b[r];
        P.rng_end[r] = rg.e[r];
        P.rng_nseg[r] = nsegs(rg.b[r], rg.e[r], pl.seg_tiles);
        nseg_total += P.rng_nseg[r];
    }
    P.lazy_wait = hs.lazy ? 1 : 0;
    P.units_per_itile = units_per_itile;
    if (nseg_total == 0) return NB200_OK;
    if (ctx->opt_deterministic) {
        const size_t need = (size_t)nseg_total * 3 * s.tpad;
        if (s.det_slots_doubles < need) {
            CK(cudaStreamSynchronize(s.compute));
            if (s.det_slots) CK(cudaFree(s.det_slots));
            s.det_slots = nullptr;
            CK(cudaMalloc(&s.det_slots, need * sizeof(double)));
            s.det_slots_doubles = need;
        }
        P.slots = s.det_slots;
        P.nseg_total = nseg_total;
    }
    const Variant& V = kVariants[pl.variant];
    ForceKernel k = pick_kernel(ctx->dim, ctx->f64, pl.variant, pl.flags);
    const int units = pl.n_itiles * nseg_total;
    const int grid = std::min(pl.grid, units);
    k<<<grid, V.block, smem_bytes(ctx->dim, ctx->f64), s.compute>>>(P);
    CK(cudaGetLastError());
    ctx->launches++;
    return NB200_OK;
}

// ---- pair-symmetric pass (FP32): see nb_force_sym.cuh.  One shard: every pair once.  Several
// shards (fused NVLink exchange attached): rank g also evaluates the blocks (g, g+off) for
// off = 1 .. floor((G-1)/2) and, for even G, half of the block against the opposite rank; the
// reaction sums on the other rank's bodies are pushed into that rank's receive slots.
// From kSymMinN bodies up the pair-symmetric pass wins even without the close-pair pre-pass (exact cut-off on every
// pair, ~10 % slower than the plain chains, but 7.5 instead of 11 FMA-pipe lane-ops per interaction and no extra launches)
constexpr size_t kSymMinN = 12288;
bool use_symmetric(const nb200_ctx* ctx, bool stepping) {
    if (ctx->opt_pos48) return true;         // the 48-bit flavour exists in the pair-symmetric kernel only (one shard)
    if (ctx->opt_symmetric == 0 || ctx->opt_deterministic) return false;
    if (ctx->opt_symmetric < 0 && !use_detect(ctx) && ctx->n < kSymMinN) return false;
    if (ctx->opt_symmetric > 0 && !use_detect(ctx) && ctx->opt_detect != 0 && ctx->n < kSymMinN) return false;   // explicit 1 keeps meaning "when the pre-pass runs" below the threshold
    if (ctx->world == 1) return true;
    // cross-rank flavour: only inside nb200_step (nb200_forces stays a rank-local call)
    return stepping && !ctx->detached && ctx->p2p_ready && ctx->exchange == 1 && ctx->world <= kMaxWorldP2P;
}

// Reaction-sum exchange of rank g of W at offset off = 1 .. floor(W/2): g pushes the sums on the bodies
// of rank g+off into slot off-1 of that rank, and finds in its own slot off-1 what rank g-off pushed.
struct SymExchange { int send_to, recv_from, slot; };
SymExchange sym_exchange(int g, int W, int off) { return SymExchange{(g + off) % W, (g - off + W) % W, off - 1}; }

int peer_index(const Shard& s, int rank) {
    for (int p = 0; p < s.n_peers; ++p)
        if (s.peer_rank[p] == rank) return p;
    return -1;
}

// The work list of rank g of G (pure host logic, also exported for the CPU tests as
// nb200_debug_sym_rows): shard g owns source tiles [g*T, (g+1)*T).
void sym_rows_for(int g, int G, int T, std::vector<NbSymRow>& rows, int tpi = NB_SYM_ITILE / NB_TILE) {
    // tpi = source tiles per i-tile
    const int lo = g * T, hi = (g + 1) * T;
    const int n_it = (T + tpi - 1) / tpi;
    rows.clear();
    for (int it = 0; it < n_it; ++it) {
        const int d0 = lo + it * tpi, d1 = std::min(d0 + tpi, hi);
        rows.push_back(NbSymRow{it, d0, d1, 0});                 // the i-tile against itself: ordered pairs
        if (d1 < hi) rows.push_back(NbSymRow{it, d1, hi, NB_ROW_SYM});
    }
    if (G > 1) {
        for (int off = 1; off <= (G - 1) / 2; ++off) {
            const int h = (g + off) % G;
            for (int it = 0; it < n_it; ++it) rows.push_back(NbSymRow{it, h * T, (h + 1) * T, NB_ROW_SYM});
        }
        if (G % 2 == 0) {
            // the block against the opposite rank is split between the two: the lower rank takes its own
            // first half of i-tiles against all of the other's bodies, the higher rank all of its targets
            // against the lower rank's remaining bodies
            const int o = (g + G / 2) % G, half = n_it / 2;
            if (g < o) {
                for (int it = 0; it < half; ++it) rows.push_back(NbSymRow{it, o * T, (o + 1) * T, NB_ROW_SYM});
            } else {
                const int b0 = std::min(o * T + half * tpi, (o + 1) * T);
                if (b0 < (o + 1) * T)
                    for (int it = 0; it < n_it; ++it) rows.push_back(NbSymRow{it, b0, (o + 1) * T, NB_ROW_SYM});
            }
        }
    }
}

// units of one row: symmetric rows are cut every seg_sub sub-tiles (subt per tile), ordered rows every seg_ord tiles
int sym_row_units(const NbSymRow& r, int subt, int seg_sub, int seg_ord) {
    const int len = r.t_end - r.t_begin;
    return (r.flags & NB_ROW_SYM) ? (len * subt + seg_sub - 1) / seg_sub : (len + seg_ord - 1) / seg_ord;
}

int build_sym_rows(nb200_ctx* ctx, Shard& s, int seg_sub, int seg_ord, int subt, bool cross, int tpi) {
    const int G = cross ? ctx->world : 1;
    const int key_seg = seg_sub * 4096 + seg_ord * 8 + subt;
    if (!s.sym_rows_host.empty() && s.sym_key_seg == key_seg && s.sym_key_world == G && s.sym_key_tpi == tpi) return NB200_OK;
    std::vector<NbSymRow>& rows = s.sym_rows_host;
    sym_rows_for(cross ? s.rank : 0, G, (int)ctx->tiles_per_shard, rows, tpi);
    s.sym_key_tpi = tpi;
    if ((int)rows.size() > s.sym_rows_cap) return fail(ctx, NB200_ESTATE, "symmetric work list overflow");
    s.sym_prefix_host.assign(rows.size() + 1, 0);
    for (size_t r = 0; r < rows.size(); ++r)
        s.sym_prefix_host[r + 1] = s.sym_prefix_host[r] + sym_row_units(rows[r], subt, seg_sub, seg_ord);
    s.sym_key_seg = key_seg;
    s.sym_key_world = G;
    CK(cudaMemcpyAsync(s.sym_rows, rows.data(), rows.size() * sizeof(NbSymRow), cudaMemcpyHostToDevice, s.compute));
    CK(cudaMemcpyAsync(s.sym_prefix, s.sym_prefix_host.data(), s.sym_prefix_host.size() * sizeof(int),
                       cudaMemcpyHostToDevice, s.compute));
    return NB200_OK;
}

// symmetric force kernel, [push of the reaction sums to their owners], finish kernel (forces or integrate)
int launch_symmetric(nb200_ctx* ctx, Shard& s, int mode, double G, double cutoff, double dt, int cur, bool with_flags,
                     const Handshake& hs = Handshake()) {
    const int D = ctx->dim;
    const bool cross = ctx->world > 1;
    const int W = ctx->world;
    const int tiles = (int)ctx->tiles_per_shard;
    // shape = targets per thread x threads per CTA (i-tile = their product); options "sym_ti", "sym_block"
    //   FP32: 8 x 128 (default: fewest hand-overs per chain), 4 x 256, 4 x 128
    //   FP64: 8 x 128 (default: two CTAs per SM at 255 registers, half the hand-overs and loads per pair), 4 x 256
    //         (one CTA per SM), 4 x 128 (three), 2 x 256 and 2 x 128 (two / four CTAs per SM under 128 registers)
    // opt-in shape for small problems on one shard: i-tiles of ONE source tile (4 targets x 64 threads, many
    // small CTAs); option "sym_itile" = 256 selects it
    const bool small = !cross && !ctx->opt_pos48 && (ctx->opt_sym_itile ? ctx->opt_sym_itile == 256 : (kSymSmallN > 0 && ctx->n < kSymSmallN));
    SymShape sh{4, 64};
    if (!small) {
        if (ctx->f64) {
            sh.ti = ctx->opt_sym_ti ? ctx->opt_sym_ti : kSymTiF64;
            sh.block = sh.ti == 8 ? 128 : ctx->opt_sym_block == 128 ? 128 : 256;
        } else {
            sh.ti = ctx->opt_sym_ti == 8 ? 8 : ctx->opt_sym_ti == 4 ? 4 : kSymTiF32;
            sh.block = sh.ti == 8 ? 128 : ctx->opt_sym_block == 128 ? 128 : 256;
        }
    }
    const int algo = ctx->opt_pos48 ? 2 : small ? 0 : std::min(ctx->opt_sym_algo >= 0 ? ctx->opt_sym_algo : kSymAlgoDefault, ctx->f64 ? 1 : 2);
    // FP64: the 8 x 128 and 4 x 128 shapes exist for the rotation only
    if (!small && ctx->f64 && algo == 0 && sh.ti >= 4) sh = SymShape{4, 256};
    const int subt = nb_sym_subtiles(ctx->f64, algo);
    // small FP32 problems (measured, profiles/r02/small_n.jsonl): 512-target i-tiles on four CTAs per SM balance better
    // up to N ~ 24576 (N=16384: 0.108 ms/step vs 0.112 with 8 x 128)
    if (!small && !ctx->f64 && !ctx->opt_sym_ti && !ctx->opt_sym_block && !cross && (long long)tiles * NB_TILE <= 24576)
        sh = SymShape{4, 128};
    // 48-bit positions: one flavour (8 x 128, decoupled rotation, general masses, exact cut-off)
    const bool hl = ctx->opt_pos48;
    if (hl) sh = SymShape{8, 128};
    const bool eqm = !hl && ctx->equal_mass && sym_has_eqm(ctx->f64, sh, algo);
    int resident = 0;
    {
        // asked once per kernel, not once per step: the query is a driver call of several microseconds (more once other
        // lazily loaded modules live in the context), and a 0.1 ms step has no host time to spare
        const void* fn = (const void*)pick_sym_kernel(D, ctx->f64, sh, algo, eqm, hl);
        if (s.occ_fn != fn) {
            int nb = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, sh.block, nb_sym_smem_bytes(D, sh.block, ctx->f64, sh.ti, algo, hl)));
            s.occ_fn = fn;
            s.occ_blocks = std::max(1, nb);
        }
        resident = s.occ_blocks * s.sms;
    }
    // work units: ~18 per resident CTA keep the tail short without paying a unit's fixed cost (target loads, first TMA
    // wait, the FP64 atomics of the target sums) too often; at most 32 tiles, at least one sub-tile
    int seg_sub = ctx->opt_seg_sub > 0 ? ctx->opt_seg_sub : ctx->opt_seg_tiles * subt;
    if (seg_sub <= 0) {
        const long long n_it0 = ((long long)tiles * NB_TILE + sh.ti * sh.block - 1) / (sh.ti * sh.block);
        const long long cells = n_it0 * (long long)ctx->ntiles / 2;     // (i-tile, source tile) cells this rank evaluates
        seg_sub = (int)std::max<long long>(1, std::min<long long>(32LL * subt, cells * subt / (18LL * resident)));
    }
    const int ti = sh.ti, block = sh.block;
    const int itile = ti * block;
    const int seg_ord = std::max(1, seg_sub / subt);
    const SymKernel kfn = pick_sym_kernel(D, ctx->f64, sh, algo, eqm, hl);
    const size_t smem = nb_sym_smem_bytes(D, block, ctx->f64, ti, algo, hl);
    if (int rc = build_sym_rows(ctx, s, seg_sub, seg_ord, subt, cross, itile / NB_TILE)) return rc;
    const int seg = seg_sub;
    NbSymParams Q;
    memset(&Q, 0, sizeof Q);
    Q.src = s.src[cur];
    Q.src_lo = hl ? s.src_lo[cur] : nullptr;
    Q.gacc = s.gacc;
    Q.gstride = (size_t)ctx->nalloc;
    Q.sched = s.sched;
    Q.rows = s.sym_rows;
    Q.row_prefix = s.sym_prefix;
    Q.suspect = with_flags ? s.suspect : nullptr;      // no pre-pass: exact cut-off on every pair
    Q.tgt_base = s.tgt_base;
    Q.own_count = tiles * NB_TILE;
    Q.n_rows = (int)s.sym_rows_host.size();
    Q.seg_sub = seg_sub;
    Q.seg_ord = seg_ord;
    Q.total_units = s.sym_prefix_host.back();
    Q.cutoff = cutoff * ctx->pos_scale * ctx->pos_scale;
    const int grid = std::min(resident, Q.total_units);
    // one shard: [pre-pass insert -> query ->] this pass -> finish kernel are chained by programmatic dependent launch
    const bool pdl = !cross && ctx->world == 1 && ctx->opt_pdl != 0;
    CK(launch_pdl(kfn, grid, block, smem, s.compute, pdl, Q));
    ctx->launches++;

    NbSymFinish F;
    memset(&F, 0, sizeof F);
    F.lo_next = (hl && mode != 0) ? s.src_lo[cur ^ 1] : nullptr;
    F.gacc = s.gacc;
    F.gstride = (size_t)ctx->nalloc;
    F.count = tiles * NB_TILE;
    F.done = s.sym_done + 1;
    if (cross) {
        const int n_ex = W / 2;                     // ranks this one sends to == ranks it receives from
        const unsigned long long seq = ++ctx->acc_seq_issued[&s - &ctx->shards[0]];
        NbSymPush U;
        memset(&U, 0, sizeof U);
        U.gacc = s.gacc;
        U.gstride = (size_t)ctx->nalloc;
        U.count = tiles * NB_TILE;
        U.seq = seq;
        U.done = s.sym_done;
        const size_t slot_doubles = (size_t)3 * tiles * NB_TILE;
        for (int off = 1; off <= n_ex; ++off) {
            const SymExchange ex = sym_exchange(s.rank, W, off);
            const int h = ex.send_to, q = ex.recv_from;
            const int ph = peer_index(s, h), pq = peer_index(s, q);
            if (ph < 0 || pq < 0) return fail(ctx, NB200_ESTATE, "peer %d/%d of rank %d is not attached", h, q, s.rank);
            const int k = U.n_dst++;
            U.body_begin[k] = (long long)h * tiles * NB_TILE;
            U.dst_slot[k] = reinterpret_cast<double*>(reinterpret_cast<char*>(s.peer_flags[ph]) + kFlagsBytes) +
                            (size_t)ex.slot * slot_doubles;
            U.dst_flag[k] = s.peer_flags[ph] + 2 * kMaxWorldP2P + s.rank;
            F.slot[F.n_src] = reinterpret_cast<const double*>(reinterpret_cast<const char*>(s.flags) + kFlagsBytes) +
                              (size_t)ex.slot * slot_doubles;
            F.flag[F.n_src] = s.flags + 2 * kMaxWorldP2P + q;
            F.sender[F.n_src] = q;
            F.n_src++;
        }
        F.seq = seq;
        const int pb = std::min(4 * s.sms, (int)(((long long)U.count * D * U.n_dst + 255) / 256));
        if (D == 3) nb_sym_push_kernel<3><<<pb, 256, 0, s.compute>>>(U);
        else nb_sym_push_kernel<2><<<pb, 256, 0, s.compute>>>(U);
        CK(cudaGetLastError());
        ctx->launches++;
    }
    NbForceParams P = base_params(ctx, s, mode, G, cutoff, dt, cur, hs);
    if (eqm) P.acc_scale *= ctx->common_mass * ctx->mass_scale;      // the pass summed without the (common) mass
    const int fb = (s.tpad + 255) / 256;
    if (ctx->f64) {
        if (D == 3) CK(launch_pdl(nb_finish_kernel<3, double>, fb, 256, 0, s.compute, pdl, P, F));
        else CK(launch_pdl(nb_finish_kernel<2, double>, fb, 256, 0, s.compute, pdl, P, F));
    } else {
        if (D == 3) CK(launch_pdl(nb_finish_kernel<3, float>, fb, 256, 0, s.compute, pdl, P, F));
        else CK(launch_pdl(nb_finish_kernel<2, float>, fb, 256, 0, s.compute, pdl, P, F));
    }
    ctx->launches++;
    if (&s == &ctx->shards[0]) {
        char buf[320];
        snprintf(buf, sizeof buf,
                 "%s: fp%d dim=%d n=%zu shards=%d pair-symmetric(TI=%d,block=%d,itile=%d) algo=%d seg=%d/%d tiles rows=%d "
                 "units=%d grid=%d tiles=%lld cutoff=%s%s + finish kernel",
                 mode ? "step" : "forces", ctx->f64 ? 64 : 32, D, ctx->n, ctx->world, ti, block, itile, algo, seg, subt, Q.n_rows, Q.total_units, grid,
                 ctx->ntiles, with_flags ? "grid-prepass(plain|exact)" : "exact",
                 cross ? " + reaction sums pushed to their owners over NVLink" : "");
        if (eqm) strncat(buf, " [equal-mass chains]", sizeof buf - strlen(buf) - 1);
        if (hl) strncat(buf, " [48-bit positions]", sizeof buf - strlen(buf) - 1);
        ctx->plan = buf;
    }
    return NB200_OK;
}

void describe_plan(nb200_ctx* ctx, const Plan& pl, const char* what) {
    char buf[256];
    const Variant& V = kVariants[pl.variant];
    snprintf(buf, sizeof buf,
             "%s: fp%d dim=%d n=%zu shards=%d variant=%d(TI=%d,JS=%d,block=%d,itile=%d) seg_tiles=%d "
             "i-tiles=%d grid=%d tiles=%lld cutoff=%s",
             what, ctx->f64 ? 64 : 32, ctx->dim, ctx->n, ctx->world, pl.variant, V.ti, V.js, V.block, V.itile(),
             pl.seg_tiles, pl.n_itiles, pl.grid, ctx->ntiles,
             pl.flags ? "grid-prepass(plain|exact)" : (ctx->f64 ? "exact" : "tracked+redo"));
    ctx->plan = buf;
}

// optional timeline of shard 0 (option "trace"): one timed event per mark, reported by nb200_plan()
int trace_mark(nb200_ctx* ctx, const Shard& s, cudaStream_t st, const char* label, int step) {
    if (!ctx->opt_trace || &s != &ctx->shards[0] || ctx->trace.size() >= 64) return NB200_OK;
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    CK(cudaEventRecord(e, st));
    ctx->trace.emplace_back(std::string(label) + std::to_string(step), e);
    return NB200_OK;
}

// split = the own tile range is segmented separately from the two remote ranges
int total_units_per_itile(const nb200_ctx* ctx, const Shard& s, const Plan& pl, bool split) {
    const int NT = (int)ctx->ntiles;
    if (!split) return nsegs(0, NT, pl.seg_tiles);
    return nsegs((int)s.tile_lo, (int)s.tile_hi, pl.seg_tiles) + nsegs(0, (int)s.tile_lo, pl.seg_tiles) +
           nsegs((int)s.tile_hi, NT, pl.seg_tiles);
}

int finish_timing(nb200_ctx* ctx) {
    double worst = 0.0;
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        CK(cudaEventSynchronize(s.ev_stop));
        CK(cudaStreamSynchronize(s.comm));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s.ev_start, s.ev_stop));
        worst = std::max(worst, (double)ms);
    }
    ctx->last_ms = worst;
    // a flag wait that ran out of time left a record: the result of this call is void and, since the ranks'
    // step counters may no longer agree, so is the context
    for (Shard& s : ctx->shards) {
        const unsigned long long w = s.err_host ? *reinterpret_cast<volatile unsigned long long*>(s.err_host) : 0ull;
        if (!w) continue;
        const int kind = (int)(w >> 56), peer = (int)((w >> 48) & 0xff);
        const unsigned long long want = w & 0xffffffffffffull;
        ctx->dead = true;
        return fail(ctx, NB200_ESTATE,
                    "rank %d: peer %d did not publish %s %llu within %ld ms (crashed, not called with the same options/steps, "
                    "or never attached); the context is unusable, destroy it on every rank",
                    s.rank, peer, kind == (int)NB_WAIT_STEP ? "step" : kind == (int)NB_WAIT_EPOCH ? "upload epoch" : "reaction-sum pass",
                    want, ctx->opt_spin_timeout_ms);
    }
    return NB200_OK;
}

int init_common(nb200_ctx* ctx) {
    ctx->acc_seq_issued.assign(ctx->shards.size(), 0ull);
    for (Shard& s : ctx->shards) {
        int rc = alloc_shard(ctx, s);
        if (rc) return rc;
    }
    return NB200_OK;
}


int publish_epoch(nb200_ctx* ctx);

// AoS staging image -> tile-planar sources (both buffers) + FP64 master state, on every shard.  With a shard-local
// image every shard packs the rows it owns and stores them into all shards' buffers (nb_pack_shard_kernel).
int pack_sources(nb200_ctx* ctx) {
    const int D = ctx->dim;
    const size_t sd = ctx->aos_stride / sizeof(double);
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        const int threads = 256;
        if (ctx->image_full) {
            const int blocks = (int)((ctx->nalloc + threads - 1) / threads);
#define NB_PACK(DD, RR)                                                                              \
    nb_pack_kernel<DD, RR><<<blocks, threads, 0, s.compute>>>(                                       \
        s.aos_dev, sd, (long long)ctx->n, ctx->nalloc, (RR*)s.src[0], (RR*)s.src[1], ctx->pos_scale, \
        ctx->mass_scale, s.tgt_base, s.tpad, s.pos, s.vel, s.mass, ctx->equal_mass ? 1 : 0)
            if (D == 3) { if (ctx->f64) NB_PACK(3, double); else NB_PACK(3, float); }
            else        { if (ctx->f64) NB_PACK(2, double); else NB_PACK(2, float); }
#undef NB_PACK
        } else {
            NbPeerBufs pb;
            memset(&pb, 0, sizeof pb);
            pb.n_peers = s.n_peers;
            for (int p = 0; p < s.n_peers; ++p) { pb.buf0[p] = s.peer_src[0][p]; pb.buf1[p] = s.peer_src[1][p]; }
            const long long nbodies = ctx->ntiles * NB_TILE;
            const int span = (int)(ctx->tiles_per_shard * NB_TILE);
            const int blocks = (int)((s.tpad + (ctx->nalloc - nbodies) + threads - 1) / threads);
#define NB_PACK(DD, RR)                                                                                              \
    nb_pack_shard_kernel<DD, RR><<<blocks, threads, 0, s.compute>>>(                                                 \
        s.aos_dev, sd, (long long)ctx->n, s.tgt_base, span, s.tpad, nbodies, ctx->nalloc, (RR*)s.src[0], (RR*)s.src[1], pb, \
        ctx->pos_scale, ctx->mass_scale, s.pos, s.vel, s.mass, ctx->equal_mass ? 1 : 0)
            if (D == 3) { if (ctx->f64) NB_PACK(3, double); else NB_PACK(3, float); }
            else        { if (ctx->f64) NB_PACK(2, double); else NB_PACK(2, float); }
#undef NB_PACK
        }
        CK(cudaGetLastError());
        ctx->launches++;
        if (ctx->opt_pos48) {
            const size_t lo_bytes = (size_t)ctx->nalloc * D * sizeof(float);
            for (int b = 0; b < 2; ++b)
                if (!s.src_lo[b]) CK(cudaMalloc(&s.src_lo[b], lo_bytes));
            const int blocks = (int)((ctx->nalloc + threads - 1) / threads);
            if (D == 3) nb_pack_lo_kernel<3><<<blocks, threads, 0, s.compute>>>(s.aos_dev, sd, (long long)ctx->n, ctx->nalloc, s.src_lo[0], s.src_lo[1], ctx->pos_scale);
            else nb_pack_lo_kernel<2><<<blocks, threads, 0, s.compute>>>(s.aos_dev, sd, (long long)ctx->n, ctx->nalloc, s.src_lo[0], s.src_lo[1], ctx->pos_scale);
            CK(cudaGetLastError());
            ctx->launches++;
        }
    }
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        CK(cudaStreamSynchronize(s.compute));
    }
    ctx->cur = 0;
    return publish_epoch(ctx);
}

// FP32 range guard: kept pairs reach 1/r'^4 <= 1/cutoff'^2, which must stay finite in FP32, so
// the scaled cut-off cutoff * ps^2 may not fall under 2^-62.  The scale is a power of two (exact).
int ensure_fp32_scale(nb200_ctx* ctx, double cutoff) {
    if (ctx->f64 || !(cutoff > 0.0) || ctx->n == 0) return NB200_OK;
    const double need = sqrt(ldexp(1.0, -62) / cutoff);
    if (ctx->pos_scale >= need) return NB200_OK;
    if (!ctx->pristine)
        return fail(ctx, NB200_ESTATE,
                    "FP32 mode: cut-off %g needs a larger source scale than the one chosen at upload; "
                    "call with this cut-off before the first step, or use NB200_FP64", cutoff);
    int ex = 0;
    frexp(need, &ex);
    ctx->pos_scale = ldexp(1.0, ex);       // power of two >= need
    return pack_sources(ctx);
}

// single-process context: map every other shard's buffers through CUDA peer access
int setup_peers_single_process(nb200_ctx* ctx) {
    const int G = (int)ctx->shards.size();
    if (G < 2 || G > kMaxWorldP2P) return NB200_OK;
    for (int a = 0; a < G; ++a)
        for (int b = 0; b < G; ++b) {
            if (a == b) continue;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, ctx->shards[a].device, ctx->shards[b].device));
            if (!can) return NB200_OK;          // stay on NCCL
        }
    for (int a = 0; a < G; ++a) {
        Shard& s = ctx->shards[a];
        CK(cudaSetDevice(s.device));
        s.n_peers = 0;
        for (int b = 0; b < G; ++b) {
            if (a == b) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(ctx->shards[b].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(ctx, NB200_ECUDA, "cudaDeviceEnablePeerAccess(%d->%d): %s", s.device, ctx->shards[b].device,
                            cudaGetErrorString(e));
            cudaGetLastError();
            const int p = s.n_peers++;
            s.peer_rank[p] = ctx->shards[b].rank;
            s.peer_src[0][p] = ctx->shards[b].src[0];
            s.peer_src[1][p] = ctx->shards[b].src[1];
            s.peer_flags[p] = ctx->shards[b].flags;
        }
    }
    ctx->p2p_ready = true;
    ctx->exchange = 1;
    return NB200_OK;
}

int ensure_nccl_single_process(nb200_ctx* ctx) {
    if (ctx->rank_mode || ctx->shards.size() < 2 || ctx->shards[0].comm_nccl) return NB200_OK;
    NcclApi* a = nccl_api();
    if (!a->error.empty()) return fail(ctx, NB200_ENCCL, "%s", a->error.c_str());
    const int G = (int)ctx->shards.size();
    std::vector<ncclComm_t> comms(G);
    std::vector<int> devs(G);
    for (int g = 0; g < G; ++g) devs[g] = ctx->shards[g].device;
    ncclResult_t r = a->CommInitAll(comms.data(), G, devs.data());
    if (r != ncclSuccess) return fail(ctx, NB200_ENCCL, "ncclCommInitAll: %s", a->GetErrorString(r));
    for (int g = 0; g < G; ++g) ctx->shards[g].comm_nccl = comms[g];
    return NB200_OK;
}

// after a (re)pack: new epoch, published to the peers once the pack kernels are done
int publish_epoch(nb200_ctx* ctx) {
    ctx->epoch++;
    ctx->epoch_base = ctx->step_index;
    if (!ctx->p2p_ready) return NB200_OK;
    for (Shard& s : ctx->shards) {
        if (s.n_peers == 0) continue;
        CK(cudaSetDevice(s.device));
        NbForceParams P;
        memset(&P, 0, sizeof P);
        P.n_peers = s.n_peers;
        P.my_rank = s.rank;
        for (int p = 0; p < s.n_peers; ++p) P.peer_flags[p] = s.peer_flags[p];
        nb_publish_epoch_kernel<<<1, 32, 0, s.compute>>>(P, kMaxWorldP2P, ctx->epoch);
        CK(cudaGetLastError());
        ctx->launches++;
        CK(cudaStreamSynchronize(s.compute));
    }
    return NB200_OK;
}

void set_fp32_scales(nb200_ctx* ctx, double xmax, double mmax, double mmin) {
    ctx->equal_mass = ctx->opt_eqm != 0 && ctx->n > 0 && mmax > 0.0 && isfinite(mmax) && mmin == mmax;
    ctx->common_mass = ctx->equal_mass ? mmax : 0.0;
    int ex = 0;
    if (xmax > 0 && isfinite(xmax)) ctx->xmax = xmax;
    if (!ctx->f64) {
        if (xmax > 0 && isfinite(xmax)) { frexp(xmax, &ex); ctx->pos_scale = ldexp(1.0, -ex); }
        if (mmax > 0 && isfinite(mmax)) { frexp(mmax, &ex); ctx->mass_scale = ldexp(1.0, -ex); }
    }
}

// Shard-local upload: bounds of the own rows on every shard, exchanged through the peers' bounds tables together
// with the "ready" handshake (no shard overwrites a peer's source buffers before that peer has finished with them);
// every rank then derives the same power-of-two scales from the same table.
int shard_local_bounds(nb200_ctx* ctx) {
    const int D = ctx->dim;
    const size_t sd = ctx->aos_stride / sizeof(double);
    const unsigned long long next_epoch = ctx->epoch + 1;
    const int ready_word = 3 * kMaxWorldP2P;
    const unsigned long long timeout_ns = (unsigned long long)std::max(1L, ctx->opt_spin_timeout_ms) * 1000000ull;
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        CK(cudaMemsetAsync(s.bounds, 0, 2 * sizeof(unsigned long long), s.compute));
        CK(cudaMemsetAsync(s.bounds + 2, 0xFF, sizeof(unsigned long long), s.compute));     // min |mass|: all ones = no body yet
        if (s.n_local > 0) {
            const int blocks = (int)std::min<long long>(4LL * s.sms, (s.n_local + 255) / 256);
            const double* rows = s.aos_dev + (size_t)s.tgt_base * sd;
            if (D == 3) nb_bounds_kernel<3><<<blocks, 256, 0, s.compute>>>(rows, sd, s.n_local, s.bounds);
            else nb_bounds_kernel<2><<<blocks, 256, 0, s.compute>>>(rows, sd, s.n_local, s.bounds);
            CK(cudaGetLastError());
            ctx->launches++;
        }
        NbForceParams P;
        memset(&P, 0, sizeof P);
        P.n_peers = s.n_peers;
        P.my_rank = s.rank;
        P.my_flags = s.flags;
        for (int p = 0; p < s.n_peers; ++p) P.peer_flags[p] = s.peer_flags[p];
        nb_publish_ready_kernel<<<1, 32, 0, s.compute>>>(P, s.bounds, kBoundsTableWord, ready_word, next_epoch);
        CK(cudaGetLastError());
        ctx->launches++;
    }
    double xmax = 0.0, mmax = 0.0, mmin = HUGE_VAL;
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        nb_wait_ready_kernel<<<1, 32, 0, s.compute>>>(s.flags, ready_word, s.n_peers, PeerRanks(s), next_epoch, timeout_ns, s.err_dev);
        CK(cudaGetLastError());
        ctx->launches++;
        double table[3 * kMaxWorldP2P];
        CK(cudaMemcpyAsync(table, s.flags + kBoundsTableWord, sizeof table, cudaMemcpyDeviceToHost, s.compute));
        CK(cudaStreamSynchronize(s.compute));
        if (s.err_host && *reinterpret_cast<volatile unsigned long long*>(s.err_host)) {
            ctx->dead = true;
            return fail(ctx, NB200_ESTATE, "rank %d: a peer did not enter upload %llu within %ld ms (uploads are collective "
                        "across the ranks of an attached exchange); the context is unusable, destroy it on every rank",
                        s.rank, next_epoch, ctx->opt_spin_timeout_ms);
        }
        for (int r = 0; r < ctx->world && r < kMaxWorldP2P; ++r) {
            if (table[3 * r] > xmax) xmax = table[3 * r];
            if (table[3 * r + 1] > mmax) mmax = table[3 * r + 1];
            if (table[3 * r + 2] < mmin) mmin = table[3 * r + 2];      // an empty shard holds all ones = NaN: never smaller
        }
    }
    set_fp32_scales(ctx, xmax, mmax, mmin);
    return NB200_OK;
}

// common tail of nb200_upload_aos / nb200_generate: the AoS image is in aos_dev on every shard
int finish_upload(nb200_ctx* ctx) {
    const int D = ctx->dim;
    const size_t sd = ctx->aos_stride / sizeof(double);
    if (!ctx->image_full) {
        if (int rc = shard_local_bounds(ctx)) return rc;
    } else if (ctx->n) {
        // bounds of the image, reduced on the device that already holds it (shard 0)
        Shard& s = ctx->shards[0];
        CK(cudaSetDevice(s.device));
        CK(cudaMemsetAsync(s.bounds, 0, 2 * sizeof(unsigned long long), s.compute));
        CK(cudaMemsetAsync(s.bounds + 2, 0xFF, sizeof(unsigned long long), s.compute));
        const int blocks = (int)std::min<long long>(4LL * s.sms, ((long long)ctx->n + 255) / 256);
        if (D == 3) nb_bounds_kernel<3><<<blocks, 256, 0, s.compute>>>(s.aos_dev, sd, (long long)ctx->n, s.bounds);
        else nb_bounds_kernel<2><<<blocks, 256, 0, s.compute>>>(s.aos_dev, sd, (long long)ctx->n, s.bounds);
        CK(cudaGetLastError());
        ctx->launches++;
        double hb[3] = {0.0, 0.0, 0.0};
        CK(cudaMemcpyAsync(hb, s.bounds, sizeof hb, cudaMemcpyDeviceToHost, s.compute));
        CK(cudaStreamSynchronize(s.compute));
        set_fp32_scales(ctx, hb[0], hb[1], hb[2]);
    } else {
        ctx->equal_mass = false;
    }
    int rc = pack_sources(ctx);
    if (rc) return rc;
    for (Shard& s : ctx->shards) s.forces_valid = false;
    ctx->pristine = true;
    ctx->cur = 0;
    ctx->uploaded = true;
    return NB200_OK;
}

}  // namespace

// =============================================================================== C ABI
// NVTX range around every public entry point that touches the device (header-only NVTX v3: a no-op until a tool such
// as Nsight Systems injects itself; SURVEY 5 "tracing").  The CUDA-event timeline of option "trace" is independent.
struct NbRange {
    explicit NbRange(const char* name) { nvtxRangePushA(name); }
    ~NbRange() { nvtxRangePop(); }
    NbRange(const NbRange&) = delete;
    NbRange& operator=(const NbRange&) = delete;
};

extern "C" {

const char* nb200_version(void) { return NB200_VERSION_STR; }

const char* nb200_last_error(const nb200_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

const char* nb200_plan(const nb200_ctx* ctx) { return ctx ? ctx->plan.c_str() : ""; }

long long nb200_launch_count(const nb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int nb200_last_elapsed_ms(const nb200_ctx* ctx, double* ms) {
    if (!ctx || !ms) return NB200_EINVAL;
    *ms = ctx->last_ms;
    return NB200_OK;
}

int nb200_get_unique_id(void* out) {
    if (!out) return fail(nullptr, NB200_EINVAL, "null unique id buffer");
    NcclApi* a = nccl_api();
    if (!a->error.empty()) return fail(nullptr, NB200_ENCCL, "%s", a->error.c_str());
    static_assert(sizeof(ncclUniqueId) == NB200_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    ncclResult_t r = a->GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, NB200_ENCCL, "ncclGetUniqueId: %s", a->GetErrorString(r));
    memcpy(out, &id, sizeof id);
    return NB200_OK;
}


int nb200_ipc_export(nb200_ctx* ctx, void* blob) {
    if (!ctx || !blob) return NB200_EINVAL;
    if (!ctx->rank_mode) return fail(ctx, NB200_ESTATE, "ipc export is for nb200_create_rank contexts");
    Shard& s = ctx->shards[0];
    CK(cudaSetDevice(s.device));
    unsigned char* out = static_cast<unsigned char*>(blob);
    memset(out, 0, NB200_IPC_BYTES);
    static_assert(3 * sizeof(cudaIpcMemHandle_t) + 8 <= NB200_IPC_BYTES, "ipc blob size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, s.src[0])); memcpy(out, &h, sizeof h);
    CK(cudaIpcGetMemHandle(&h, s.src[1])); memcpy(out + sizeof h, &h, sizeof h);
    CK(cudaIpcGetMemHandle(&h, s.flags));  memcpy(out + 2 * sizeof h, &h, sizeof h);
    const int rank = s.rank, world = ctx->world;
    memcpy(out + 3 * sizeof h, &rank, 4);
    memcpy(out + 3 * sizeof h + 4, &world, 4);
    return NB200_OK;
}

int nb200_ipc_attach(nb200_ctx* ctx, const void* blobs, int count) {
    NbRange nvtx_range("nb200_ipc_attach");
    if (!ctx || !blobs) return NB200_EINVAL;
    if (!ctx->rank_mode) return fail(ctx, NB200_ESTATE, "ipc attach is for nb200_create_rank contexts");
    if (count != ctx->world) return fail(ctx, NB200_EINVAL, "expected %d blobs, got %d", ctx->world, count);
    if (ctx->world > kMaxWorldP2P) return fail(ctx, NB200_EINVAL, "peer-store exchange supports up to %d GPUs", kMaxWorldP2P);
    if (ctx->uploaded && !ctx->pristine) return fail(ctx, NB200_ESTATE, "attach before the first step");
    Shard& s = ctx->shards[0];
    if (s.ipc_opened || ctx->p2p_ready) return fail(ctx, NB200_ESTATE, "the peer-store exchange is already attached");
    CK(cudaSetDevice(s.device));
    const unsigned char* in = static_cast<const unsigned char*>(blobs);
    s.n_peers = 0;
    // on any failure the mappings opened so far are closed again: the context stays on its NCCL communicator
    auto undo = [&](void* a0, void* a1) {
        if (a0) cudaIpcCloseMemHandle(a0);
        if (a1) cudaIpcCloseMemHandle(a1);
        for (int p = 0; p < s.n_peers; ++p) {
            cudaIpcCloseMemHandle(s.peer_src[0][p]);
            cudaIpcCloseMemHandle(s.peer_src[1][p]);
            cudaIpcCloseMemHandle(s.peer_flags[p]);
        }
        s.n_peers = 0;
        cudaGetLastError();
    };
    for (int r = 0; r < count; ++r) {
        const unsigned char* b = in + (size_t)r * NB200_IPC_BYTES;
        int brank = -1, bworld = -1;
        memcpy(&brank, b + 3 * sizeof(cudaIpcMemHandle_t), 4);
        memcpy(&bworld, b + 3 * sizeof(cudaIpcMemHandle_t) + 4, 4);
        if (brank != r || bworld != ctx->world) {
            undo(nullptr, nullptr);
            return fail(ctx, NB200_EINVAL, "blob %d is from rank %d of %d", r, brank, bworld);
        }
        if (r == s.rank) continue;
        cudaIpcMemHandle_t h;
        void *m0 = nullptr, *m1 = nullptr, *mf = nullptr;
        memcpy(&h, b, sizeof h);
        cudaError_t e = cudaIpcOpenMemHandle(&m0, h, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) {
            memcpy(&h, b + sizeof h, sizeof h);
            e = cudaIpcOpenMemHandle(&m1, h, cudaIpcMemLazyEnablePeerAccess);
        }
        if (e == cudaSuccess) {
            memcpy(&h, b + 2 * sizeof h, sizeof h);
            e = cudaIpcOpenMemHandle(&mf, h, cudaIpcMemLazyEnablePeerAccess);
        }
        if (e != cudaSuccess) {
            undo(m0, m1);
            return fail(ctx, NB200_ECUDA, "cudaIpcOpenMemHandle of rank %d's buffers failed: %s", r, cudaGetErrorString(e));
        }
        const int p = s.n_peers++;
        s.peer_src[0][p] = m0;
        s.peer_src[1][p] = m1;
        s.peer_flags[p] = static_cast<unsigned long long*>(mf);
        s.peer_rank[p] = r;
    }
    s.ipc_opened = true;
    ctx->p2p_ready = true;
    ctx->exchange = 1;
    // the epoch of an earlier upload was not published to anybody: publish it now
    if (ctx->uploaded) { ctx->epoch--; return publish_epoch(ctx); }
    return NB200_OK;
}

int nb200_create(nb200_ctx** out, int dim, size_t n, int precision, int ngpus) {
    if (!out) return fail(nullptr, NB200_EINVAL, "null out pointer");
    *out = nullptr;
    int have = 0;
    cudaError_t e = cudaGetDeviceCount(&have);
    if (e != cudaSuccess || have == 0)
        return fail(nullptr, NB200_ECUDA, "no CUDA device: %s (libnb200 has no CPU fallback)",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (ngpus < 1 || ngpus > have) return fail(nullptr, NB200_EINVAL, "ngpus=%d but %d device(s) visible", ngpus, have);
    nb200_ctx* ctx = new nb200_ctx();
    int rc = layout(ctx, dim, n, precision, ngpus);
    if (rc) { delete ctx; return rc; }
    std::vector<int> devs;
    if (const char* env = getenv("NB200_DEVICES")) {
        for (const char* p = env; *p && (int)devs.size() < ngpus;) {
            devs.push_back(atoi(p));
            p = strchr(p, ',');
            if (!p) break;
            ++p;
        }
    }
    for (int i = (int)devs.size(); i < ngpus; ++i) devs.push_back(i);
    ctx->shards.resize(ngpus);
    for (int g = 0; g < ngpus; ++g) place_shard(ctx, ctx->shards[g], g, devs[g]);
    rc = init_common(ctx);
    if (rc == NB200_OK && ngpus > 1) {
        // preferred exchange: peer stores fused into the epilogue; NCCL (lazily initialised) otherwise
        rc = setup_peers_single_process(ctx);
        if (rc == NB200_OK && !ctx->p2p_ready) rc = ensure_nccl_single_process(ctx);
    }
    if (rc) {
        g_create_error = ctx->error;
        nb200_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return NB200_OK;
}

int nb200_create_rank(nb200_ctx** out, int dim, size_t n, int precision, int device, int rank, int world,
                      const void* unique_id) {
    if (!out) return fail(nullptr, NB200_EINVAL, "null out pointer");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(nullptr, NB200_EINVAL, "bad rank %d / world %d", rank, world);
    // unique_id == NULL with world > 1 builds a DETACHED shard (no communicator): forces() works,
    // step() is limited to nsteps == 1 (the other shards' rows of the next buffer are not refreshed;
    // re-upload before stepping again).  This is the "virtual rank" hook SURVEY.md section 4 asks for so
    // that the sharded passes can be exercised on one GPU.
    int have = 0;
    cudaError_t e = cudaGetDeviceCount(&have);
    if (e != cudaSuccess || have == 0)
        return fail(nullptr, NB200_ECUDA, "no CUDA device: %s (libnb200 has no CPU fallback)",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= have) return fail(nullptr, NB200_EINVAL, "device %d not visible (%d devices)", device, have);
    nb200_ctx* ctx = new nb200_ctx();
    int rc = layout(ctx, dim, n, precision, world);
    if (rc) { delete ctx; return rc; }
    ctx->rank_mode = true;
    ctx->shards.resize(1);
    place_shard(ctx, ctx->shards[0], rank, device);
    rc = init_common(ctx);
    ctx->detached = world > 1 && !unique_id;
    if (rc == NB200_OK && world > 1 && unique_id) {
        NcclApi* a = nccl_api();
        if (!a->error.empty()) rc = fail(ctx, NB200_ENCCL, "%s", a->error.c_str());
        else {
            ncclUniqueId id;
            memcpy(&id, unique_id, sizeof id);
            cudaSetDevice(device);
            ncclResult_t r = a->CommInitRank(&ctx->shards[0].comm_nccl, world, id, rank);
            if (r != ncclSuccess) rc = fail(ctx, NB200_ENCCL, "ncclCommInitRank: %s", a->GetErrorString(r));
        }
    }
    if (rc) {
        g_create_error = ctx->error;
        nb200_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return NB200_OK;
}

void nb200_destroy(nb200_ctx* ctx) {
    if (!ctx) return;
    for (Shard& s : ctx->shards) free_shard(s);
    delete ctx;
}

int nb200_shard_range(const nb200_ctx* ctx, size_t* lo, size_t* hi) {
    if (!ctx || !lo || !hi) return NB200_EINVAL;
    if (!ctx->rank_mode) { *lo = 0; *hi = ctx->n; return NB200_OK; }
    const Shard& s = ctx->shards[0];
    *lo = (size_t)std::min<long long>(s.tgt_base, (long long)ctx->n);
    *hi = *lo + (size_t)s.n_local;
    return NB200_OK;
}

int nb200_set_option(nb200_ctx* ctx, const char* key, long value) {
    if (!ctx || !key) return NB200_EINVAL;
    if (!strcmp(key, "variant")) ctx->opt_variant = (value >= 0 && value < kNumVariants) ? (int)value : -1;
    else if (!strcmp(key, "seg_tiles")) ctx->opt_seg_tiles = (int)std::max(0L, value);
    else if (!strcmp(key, "seg_sub")) ctx->opt_seg_sub = (int)std::max(0L, value);
    else if (!strcmp(key, "equal_mass")) {
        // 0 switches the equal-mass flavour of the pair kernels off; the padding layout is fixed at upload, so set it before
        if (ctx->uploaded && (value != 0) != (ctx->opt_eqm != 0))
            return fail(ctx, NB200_ESTATE, "set 'equal_mass' before the upload (it decides where the padding bodies go)");
        ctx->opt_eqm = value < 0 ? -1 : (value != 0);
    }
    else if (!strcmp(key, "deterministic")) {
        if (value != 0 && ctx->opt_pos48) return fail(ctx, NB200_EINVAL, "'deterministic' and 'fp32_positions' = 48 exclude each other");
        ctx->opt_deterministic = value != 0;
    }
    else if (!strcmp(key, "fp32_positions")) {
        // 48: FP32 pair arithmetic on positions kept as float pairs (hi + lo); 24: the plain FP32 mode
        if (value != 24 && value != 48) return fail(ctx, NB200_EINVAL, "fp32_positions is 24 or 48");
        if (value == 48) {
            if (ctx->f64) return fail(ctx, NB200_EINVAL, "fp32_positions applies to NB200_FP32 contexts");
            if (ctx->world != 1 || ctx->shards.size() != 1) return fail(ctx, NB200_EINVAL, "fp32_positions = 48 runs on one shard (the lo rows are not exchanged)");
            if (ctx->opt_deterministic) return fail(ctx, NB200_EINVAL, "'deterministic' and 'fp32_positions' = 48 exclude each other");
        }
        if (ctx->uploaded && (value == 48) != ctx->opt_pos48)
            return fail(ctx, NB200_ESTATE, "set 'fp32_positions' before the upload (the lo rows are written by the pack kernel)");
        ctx->opt_pos48 = value == 48;
    }
    else if (!strcmp(key, "pdl")) ctx->opt_pdl = value < 0 ? -1 : (value != 0);
    else if (!strcmp(key, "shard_upload")) ctx->opt_shard_upload = value < 0 ? -1 : (value != 0);
    else if (!strcmp(key, "grid_mult")) ctx->opt_grid_mult = (int)std::max(0L, value);
    else if (!strcmp(key, "overlap")) ctx->opt_overlap = value < 0 ? -1 : (value != 0);
    else if (!strcmp(key, "trace")) ctx->opt_trace = value != 0;
    else if (!strcmp(key, "detect")) ctx->opt_detect = value < 0 ? -1 : (value != 0);
    else if (!strcmp(key, "symmetric")) ctx->opt_symmetric = value < 0 ? -1 : (value != 0);
    else if (!strcmp(key, "sym_ti")) ctx->opt_sym_ti = (value == 8 || value == 4 || value == 2) ? (int)value : 0;
    else if (!strcmp(key, "spin_timeout_ms")) ctx->opt_spin_timeout_ms = value > 0 ? value : 30000;
    else if (!strcmp(key, "debug_fake_peer")) {
        // test hook for the failure path (one GPU, no second process): a detached shard treats its OWN buffers as the
        // buffers of its neighbour rank, whose flags nobody ever writes -- every wait on that peer must time out
        if (!ctx->rank_mode || !ctx->detached || ctx->world < 2)
            return fail(ctx, NB200_ESTATE, "debug_fake_peer needs a detached rank context (world > 1, no unique id)");
        Shard& s = ctx->shards[0];
        s.n_peers = 1;
        s.peer_rank[0] = (s.rank + 1) % ctx->world;
        s.peer_src[0][0] = s.src[0];
        s.peer_src[1][0] = s.src[1];
        s.peer_flags[0] = s.flags;
        ctx->detached = false;
        ctx->p2p_ready = true;
        ctx->exchange = 1;
    }
    else if (!strcmp(key, "sym_block")) ctx->opt_sym_block = (value == 128 || value == 256) ? (int)value : 0;
    else if (!strcmp(key, "sym_algo")) ctx->opt_sym_algo = (value >= 0 && value <= 2) ? (int)value : -1;
    else if (!strcmp(key, "sym_itile")) ctx->opt_sym_itile = value == 256 ? 256 : value == 1024 ? 1024 : 0;
    else if (!strcmp(key, "exchange")) {
        if (value == 1 && !ctx->p2p_ready) return fail(ctx, NB200_ESTATE, "peer-store exchange is not attached");
        if (value == 0 && ctx->rank_mode && ctx->world > 1 && !ctx->detached && !ctx->shards[0].comm_nccl)
            return fail(ctx, NB200_ESTATE, "no NCCL communicator in this context");
        ctx->exchange = value != 0;
    }
    else return fail(ctx, NB200_EINVAL, "unknown option '%s'", key);
    return NB200_OK;
}

int nb200_upload_aos(nb200_ctx* ctx, const void* bodies, size_t stride) {
    NbRange nvtx_range("nb200_upload_aos");
    if (!ctx) return NB200_EINVAL;
    if (ctx->dead) return fail(ctx, NB200_ESTATE, "context is unusable after a peer handshake timeout: destroy it");
    const int D = ctx->dim;
    const size_t min_stride = (size_t)(2 * D + 1) * sizeof(double);
    if (ctx->n && !bodies) return fail(ctx, NB200_EINVAL, "null bodies");
    if (stride < min_stride || stride % sizeof(double))
        return fail(ctx, NB200_EINVAL, "stride %zu: Body<%d> needs >= %zu bytes, multiple of 8", stride, D, min_stride);
    ctx->aos_stride = stride;
    // FP32 pair math runs on power-of-two-scaled sources (exact): |x'| <= 1, m' <= 1
    ctx->pos_scale = ctx->mass_scale = 1.0;
    ctx->xmax = 1.0;
    // With the peer-store exchange attached every shard takes only the rows it owns from the caller's array and stores
    // their source rows into all shards' buffers itself ("shard_upload" option, default on): 1/G of the PCIe traffic and of
    // the packing per GPU.  The call is then collective over the ranks (bounded wait, NB200_ESTATE if a peer never comes).
    ctx->image_full = !(ctx->world > 1 && !ctx->detached && ctx->p2p_ready && ctx->exchange == 1 && ctx->opt_shard_upload != 0);
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        const size_t bytes = std::max<size_t>(1, ctx->n) * stride;
        if (s.aos_dev && s.aos_bytes < bytes) { CK(cudaFree(s.aos_dev)); s.aos_dev = nullptr; }
        if (!s.aos_dev) { CK(cudaMalloc(&s.aos_dev, bytes)); s.aos_bytes = bytes; }
        if (ctx->image_full) {
            if (ctx->n) CK(cudaMemcpyAsync(s.aos_dev, bodies, ctx->n * stride, cudaMemcpyHostToDevice, s.compute));
        } else if (s.n_local > 0) {
            const size_t off = (size_t)s.tgt_base * stride;
            CK(cudaMemcpyAsync(reinterpret_cast<char*>(s.aos_dev) + off, static_cast<const char*>(bodies) + off,
                               (size_t)s.n_local * stride, cudaMemcpyHostToDevice, s.compute));
        }
    }
    return finish_upload(ctx);
}

int nb200_generate(nb200_ctx* ctx, int kind, unsigned long long seed, double G) {
    NbRange nvtx_range("nb200_generate");
    if (!ctx) return NB200_EINVAL;
    if (ctx->dead) return fail(ctx, NB200_ESTATE, "context is unusable after a peer handshake timeout: destroy it");
    const int D = ctx->dim;
    if (kind < 0 || kind > 2) return fail(ctx, NB200_EINVAL, "generator kind %d: 0 reference range, 1 uniform cube, 2 Plummer", kind);
    if (kind == 2 && D != 3) return fail(ctx, NB200_EINVAL, "the Plummer generator is 3D only");
    if (!(G > 0.0)) return fail(ctx, NB200_EINVAL, "G must be positive");
    const size_t stride = (size_t)(2 * D + 1) * sizeof(double);
    ctx->aos_stride = stride;
    const size_t sd = stride / sizeof(double);
    ctx->pos_scale = ctx->mass_scale = 1.0;
    ctx->xmax = 1.0;
    ctx->image_full = true;              // every shard generates (and packs) all n bodies: no PCIe, no exchange
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        const size_t bytes = std::max<size_t>(1, ctx->n) * stride;
        if (s.aos_dev && s.aos_bytes < bytes) { CK(cudaFree(s.aos_dev)); s.aos_dev = nullptr; }
        if (!s.aos_dev) { CK(cudaMalloc(&s.aos_dev, bytes)); s.aos_bytes = bytes; }
        if (ctx->n) {
            const int blocks = (int)((ctx->n + 255) / 256);
            if (D == 3) nb_generate_kernel<3><<<blocks, 256, 0, s.compute>>>(s.aos_dev, sd, (long long)ctx->n, kind, seed, G);
            else nb_generate_kernel<2><<<blocks, 256, 0, s.compute>>>(s.aos_dev, sd, (long long)ctx->n, kind, seed, G);
            CK(cudaGetLastError());
            ctx->launches++;
        }
    }
    return finish_upload(ctx);
}

int nb200_download_aos(nb200_ctx* ctx, void* bodies, size_t stride) {
    NbRange nvtx_range("nb200_download_aos");
    if (!ctx) return NB200_EINVAL;
    if (!ctx->uploaded) return fail(ctx, NB200_ESTATE, "download before upload");
    const int D = ctx->dim;
    if (stride < (size_t)(2 * D + 1) * sizeof(double) || stride % sizeof(double))
        return fail(ctx, NB200_EINVAL, "bad stride %zu", stride);
    if (ctx->n && !bodies) return fail(ctx, NB200_EINVAL, "null bodies");
    const size_t sd = stride / sizeof(double);
    for (Shard& s : ctx->shards) {
        if (s.n_local <= 0) continue;
        CK(cudaSetDevice(s.device));
        // reuse the staging image: rows [tgt_base, tgt_base+n_local) hold pos/vel, mass as uploaded
        if (stride != ctx->aos_stride) return fail(ctx, NB200_EINVAL, "download stride %zu != upload stride %zu", stride, ctx->aos_stride);
        double* rows = s.aos_dev + (size_t)s.tgt_base * sd;
        const int threads = 256;
        const int blocks = (int)((s.n_local + threads - 1) / threads);
        if (D == 3) nb_unpack_kernel<3><<<blocks, threads, 0, s.compute>>>(rows, sd, s.n_local, s.tpad, s.pos, s.vel);
        else nb_unpack_kernel<2><<<blocks, threads, 0, s.compute>>>(rows, sd, s.n_local, s.tpad, s.pos, s.vel);
        CK(cudaGetLastError());
        ctx->launches++;
        char* dst = static_cast<char*>(bodies) + (size_t)s.tgt_base * stride;
        if (stride == (size_t)(2 * D + 1) * sizeof(double)) {
            // packed Body<D> records: one contiguous copy of whole rows (the mass column of the image
            // still holds the uploaded masses, so the caller gets back the bytes it passed in)
            CK(cudaMemcpyAsync(dst, rows, (size_t)s.n_local * stride, cudaMemcpyDeviceToHost, s.compute));
        } else {
            // padded records: position+velocity only, the caller's padding stays untouched
            CK(cudaMemcpy2DAsync(dst, stride, rows, stride, (size_t)2 * D * sizeof(double), (size_t)s.n_local,
                                 cudaMemcpyDeviceToHost, s.compute));
        }
    }
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        CK(cudaStreamSynchronize(s.compute));
    }
    return NB200_OK;
}

// Vector<D>::normalized() zeroes the direction of any pair with r < 1e-10 (vector.h:95), so a
// cut-off below 1e-20 on r^2 cannot be expressed by the reference: clamp to that guard.
static inline double effective_cutoff(double c) { return c > 1e-20 ? c : 1e-20; }

int nb200_forces(nb200_ctx* ctx, double G, double cutoff_r2, double* forces_out) {
    NbRange nvtx_range("nb200_forces");
    if (!ctx) return NB200_EINVAL;
    if (ctx->dead) return fail(ctx, NB200_ESTATE, "context is unusable after a peer handshake timeout: destroy it");
    cutoff_r2 = effective_cutoff(cutoff_r2);
    if (!ctx->uploaded) return fail(ctx, NB200_ESTATE, "forces before upload");
    if (ctx->n && !forces_out) return fail(ctx, NB200_EINVAL, "null forces_out");
    const int D = ctx->dim;
    if (int rc0 = ensure_fp32_scale(ctx, cutoff_r2)) return rc0;
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        Plan pl;
        int rc = make_plan(ctx, s, &pl);
        if (rc) return rc;
        if (&s == &ctx->shards[0]) describe_plan(ctx, pl, "forces");
        CK(cudaStreamWaitEvent(s.compute, s.ev_gather[ctx->cur], 0));
        CK(cudaEventRecord(s.ev_start, s.compute));
        if (pl.flags) { if (int rcd = launch_detect(ctx, s, cutoff_r2, ctx->cur)) return rcd; }
        if (use_symmetric(ctx, false)) {
            rc = launch_symmetric(ctx, s, 0, G, cutoff_r2, 0.0, ctx->cur, pl.flags);
        } else {
            Ranges all(0, (int)ctx->ntiles);
            rc = launch_pass(ctx, s, pl, all, (unsigned)total_units_per_itile(ctx, s, pl, false), 0, G, cutoff_r2, 0.0, ctx->cur);
        }
        if (rc) return rc;
        CK(cudaEventRecord(s.ev_stop, s.compute));
        s.forces_valid = true;
        if (s.n_local > 0)
            CK(cudaMemcpyAsync(forces_out + (size_t)s.tgt_base * D, s.forces, (size_t)s.n_local * D * sizeof(double),
                               cudaMemcpyDeviceToHost, s.compute));
    }
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        CK(cudaStreamSynchronize(s.compute));
    }
    return finish_timing(ctx);
}

int nb200_step(nb200_ctx* ctx, double G, double cutoff_r2, double dt, int nsteps) {
    NbRange nvtx_range("nb200_step");
    if (!ctx) return NB200_EINVAL;
    if (ctx->dead) return fail(ctx, NB200_ESTATE, "context is unusable after a peer handshake timeout: destroy it");
    cutoff_r2 = effective_cutoff(cutoff_r2);
    if (!ctx->uploaded) return fail(ctx, NB200_ESTATE, "step before upload");
    if (nsteps < 0) return fail(ctx, NB200_EINVAL, "nsteps < 0");
    if (ctx->detached && nsteps > 1)
        return fail(ctx, NB200_ESTATE, "detached shard (no communicator): only nsteps == 1 is defined");
    if (int rc0 = ensure_fp32_scale(ctx, cutoff_r2)) return rc0;
    if (nsteps > 0) ctx->pristine = false;
    const bool multi = ctx->world > 1 && !ctx->detached;
    const bool p2p = multi && ctx->exchange == 1;          // rows pushed by the epilogue over NVLink
    const bool use_nccl = multi && !p2p;
    // local|remote split: with NCCL it hides the all-gather behind the local pass; with the fused
    // peer-store exchange it lets a rank start on its own sources before its peers finish.
    // auto: own-rows-first for the fused exchange and for detached shards; no split around NCCL (its
    // kernels spin on SMs next to our persistent CTAs -- measured slower than exposing the gather)
    const bool split = ctx->world > 1 && !ctx->opt_deterministic && (ctx->opt_overlap < 0 ? !use_nccl : ctx->opt_overlap != 0);
    if (use_nccl) { if (int rcn = ensure_nccl_single_process(ctx)) return rcn; }
    NcclApi* nccl = use_nccl ? nccl_api() : nullptr;
    if (use_nccl && !ctx->shards[0].comm_nccl) return fail(ctx, NB200_ESTATE, "no exchange attached (NCCL or peer stores)");
    std::vector<Plan> plans(ctx->shards.size());
    for (size_t i = 0; i < ctx->shards.size(); ++i) {
        CK(cudaSetDevice(ctx->shards[i].device));
        int rc = make_plan(ctx, ctx->shards[i], &plans[i]);
        if (rc) return rc;
    }
    describe_plan(ctx, plans[0], !multi ? (split ? "step(local|remote, detached)" : "step")
                                 : p2p ? (split ? "step(own rows first, fused NVLink peer stores)" : "step(fused NVLink peer stores)")
                                       : (split ? "step(local|ncclAllGather|remote)" : "step(ncclAllGather)"));
    for (auto& t : ctx->trace) cudaEventDestroy(t.second);
    ctx->trace.clear();
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        CK(cudaEventRecord(s.ev_start, s.compute));
    }
    const int NT = (int)ctx->ntiles;
    const size_t rs = ctx->f64 ? 8 : 4;
    const size_t shard_elems = (size_t)ctx->tiles_per_shard * NB_TILE * (ctx->dim + 1);
    for (int step = 0; step < nsteps; ++step) {
        const int cur = ctx->cur, nxt = cur ^ 1;
        const unsigned long long k = ++ctx->step_index;
        Handshake hs_remote;                   // for the pass that reads remote rows and runs the epilogues
        if (p2p) {
            hs_remote.exchange = true;
            hs_remote.wait_step = (k - 1 > ctx->epoch_base) ? k - 1 : 0;   // rows of step k-1, unless they are the upload
            hs_remote.wait_epoch = ctx->epoch;                             // every peer finished (re)packing its buffers
            hs_remote.signal_step = k;
        }
        Handshake hs_local;
        hs_local.exchange = p2p;
        for (size_t i = 0; i < ctx->shards.size(); ++i) {
            Shard& s = ctx->shards[i];
            const Plan& pl = plans[i];
            CK(cudaSetDevice(s.device));
            const unsigned upi = (unsigned)total_units_per_itile(ctx, s, pl, split);
            int rc;
            const bool symmetric = use_symmetric(ctx, true);
            if (pl.flags || symmetric) {
                // the pre-pass and the pair-symmetric pass read every source row of this step: take the exchange
                // handshake first
                if (use_nccl) CK(cudaStreamWaitEvent(s.compute, s.ev_gather[cur], 0));
                if (p2p && s.n_peers > 0 && (hs_remote.wait_step | hs_remote.wait_epoch)) {
                    nb_wait_flags_kernel<<<1, 32, 0, s.compute>>>(s.flags, kMaxWorldP2P, s.n_peers, PeerRanks(s),
                                                                  hs_remote.wait_step, hs_remote.wait_epoch,
                                                                  (unsigned long long)std::max(1L, ctx->opt_spin_timeout_ms) * 1000000ull, s.err_dev);
                    CK(cudaGetLastError());
                    ctx->launches++;
                }
                if (pl.flags) {
                    rc = launch_detect(ctx, s, cutoff_r2, cur);
                    if (rc) return rc;
                }
            }
            if (symmetric) {
                rc = launch_symmetric(ctx, s, 1, G, cutoff_r2, dt, cur, pl.flags, hs_remote);
            } else if (split && use_nccl) {
                // two launches around the all-gather event: own sources, then the other shards'
                Ranges own((int)s.tile_lo, (int)s.tile_hi);
                trace_mark(ctx, s, s.compute, "A>", step);
                rc = launch_pass(ctx, s, pl, own, upi, 1, G, cutoff_r2, dt, cur, hs_local);
                if (rc) return rc;
                trace_mark(ctx, s, s.compute, "A<", step);
                CK(cudaStreamWaitEvent(s.compute, s.ev_gather[cur], 0));
                trace_mark(ctx, s, s.compute, "B>", step);
                Ranges rest(0, (int)s.tile_lo, (int)s.tile_hi, NT);
                rc = launch_pass(ctx, s, pl, rest, upi, 1, G, cutoff_r2, dt, cur, hs_remote);
            } else if (split) {
                // ONE launch, own rows first; a CTA takes the peer handshake only when it reaches its
                // first remote unit, so a rank that is ahead keeps computing on its own rows
                Ranges own_first((int)s.tile_lo, (int)s.tile_hi, 0, (int)s.tile_lo, (int)s.tile_hi, NT);
                Handshake hs = hs_remote;
                hs.lazy = true;
                rc = launch_pass(ctx, s, pl, own_first, upi, 1, G, cutoff_r2, dt, cur, hs);
            } else {
                if (use_nccl) CK(cudaStreamWaitEvent(s.compute, s.ev_gather[cur], 0));
                Ranges all(0, NT);
                rc = launch_pass(ctx, s, pl, all, upi, 1, G, cutoff_r2, dt, cur, hs_remote);
            }
            if (rc) return rc;
            trace_mark(ctx, s, s.compute, "B<", step);
            if (use_nccl) {
                CK(cudaEventRecord(s.ev_pass_done, s.compute));
                CK(cudaStreamWaitEvent(s.comm, s.ev_pass_done, 0));
            }
        }
        if (use_nccl) {
            // in-place all-gather of the freshly integrated shards into the next source buffer
            if (ctx->shards.size() > 1) CKN(nccl->GroupStart());
            for (Shard& s : ctx->shards) {
                CK(cudaSetDevice(s.device));
                char* base = static_cast<char*>(s.src[nxt]);
                trace_mark(ctx, s, s.comm, "G>", step);
                CKN(nccl->AllGather(base + (size_t)s.rank * shard_elems * rs, base, shard_elems,
                                    ctx->f64 ? ncclDouble : ncclFloat, s.comm_nccl, s.comm));
            }
            if (ctx->shards.size() > 1) CKN(nccl->GroupEnd());
            for (Shard& s : ctx->shards) {
                CK(cudaSetDevice(s.device));
                CK(cudaEventRecord(s.ev_gather[nxt], s.comm));
                trace_mark(ctx, s, s.comm, "G<", step);
            }
        }
        ctx->cur = nxt;
    }
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        if (use_nccl) CK(cudaStreamWaitEvent(s.compute, s.ev_gather[ctx->cur], 0));
        if (p2p && nsteps > 0 && s.n_peers > 0) {
            // the call returns (and its device time ends) only when every peer has published the last
            // step: our source buffer is complete and no peer store into it is still in flight
            nb_wait_flags_kernel<<<1, 32, 0, s.compute>>>(s.flags, kMaxWorldP2P, s.n_peers, PeerRanks(s), ctx->step_index, 0ull,
                                                          (unsigned long long)std::max(1L, ctx->opt_spin_timeout_ms) * 1000000ull, s.err_dev);
            CK(cudaGetLastError());
            ctx->launches++;
            trace_mark(ctx, s, s.compute, "W<", nsteps - 1);
        }
        CK(cudaEventRecord(s.ev_stop, s.compute));
    }
    int rc_t = finish_timing(ctx);
    if (rc_t == NB200_OK && !ctx->trace.empty()) {
        CK(cudaSetDevice(ctx->shards[0].device));
        ctx->plan += " | trace(ms):";
        for (auto& t : ctx->trace) {
            float ms = 0.f;
            cudaEventSynchronize(t.second);
            cudaEventElapsedTime(&ms, ctx->shards[0].ev_start, t.second);
            char buf[48];
            snprintf(buf, sizeof buf, " %s=%.3f", t.first.c_str(), ms);
            ctx->plan += buf;
        }
    }
    return rc_t;
}

int nb200_energy(nb200_ctx* ctx, double G, double cutoff_r2, double* kinetic, double* potential) {
    NbRange nvtx_range("nb200_energy");
    if (!ctx || !kinetic || !potential) return NB200_EINVAL;
    if (!ctx->uploaded) return fail(ctx, NB200_ESTATE, "energy before upload");
    cutoff_r2 = effective_cutoff(cutoff_r2);
    double ke = 0.0, pe = 0.0;
    const double cs = cutoff_r2 * ctx->pos_scale * ctx->pos_scale;
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        CK(cudaStreamWaitEvent(s.compute, s.ev_gather[ctx->cur], 0));
        CK(cudaMemsetAsync(s.energy, 0, 2 * sizeof(double), s.compute));
        if (s.n_local > 0) {
            const int threads = 256;
            const int blocks = (int)((s.n_local + threads * NB_ENERGY_TI - 1) / (threads * NB_ENERGY_TI));
#define NB_EN(DD, RR)                                                                                   \
    nb_energy_kernel<DD, RR><<<blocks, threads, 0, s.compute>>>(                                         \
        (const RR*)s.src[ctx->cur], ctx->ntiles, s.tgt_base, s.n_local, s.tpad, s.vel, s.mass, G, cs,    \
        1.0 / ctx->pos_scale, 1.0 / ctx->mass_scale, s.energy)
            if (ctx->dim == 3) { if (ctx->f64) NB_EN(3, double); else NB_EN(3, float); }
            else               { if (ctx->f64) NB_EN(2, double); else NB_EN(2, float); }
#undef NB_EN
            CK(cudaGetLastError());
            ctx->launches++;
        }
    }
    for (Shard& s : ctx->shards) {
        CK(cudaSetDevice(s.device));
        double h[2];
        CK(cudaMemcpyAsync(h, s.energy, sizeof h, cudaMemcpyDeviceToHost, s.compute));
        CK(cudaStreamSynchronize(s.compute));
        ke += h[0];
        pe += h[1];
    }
    *kinetic = ke;
    *potential = pe;
    return NB200_OK;
}

int nb200_debug_sym_rows(size_t n, int world, int rank, int* rows_out, int cap) {
    if (world < 1 || rank < 0 || rank >= world || (cap > 0 && !rows_out)) return NB200_EINVAL;
    const long long tiles = std::max<long long>(1, ((long long)n + NB_TILE - 1) / NB_TILE);
    const int T = (int)((tiles + world - 1) / world);
    std::vector<NbSymRow> rows;
    sym_rows_for(rank, world, T, rows);
    for (int r = 0; r < (int)rows.size() && r < cap; ++r) {
        rows_out[4 * r + 0] = rows[r].it;
        rows_out[4 * r + 1] = rows[r].t_begin;
        rows_out[4 * r + 2] = rows[r].t_end;
        rows_out[4 * r + 3] = rows[r].flags;
    }
    return (int)rows.size();
}

int nb200_debug_sym_exchange(int world, int rank, int* out, int cap) {
    if (world < 1 || rank < 0 || rank >= world || (cap > 0 && !out)) return NB200_EINVAL;
    const int n_ex = world / 2;
    for (int off = 1; off <= n_ex && off <= cap; ++off) {
        const SymExchange ex = sym_exchange(rank, world, off);
        out[3 * (off - 1) + 0] = ex.send_to;
        out[3 * (off - 1) + 1] = ex.recv_from;
        out[3 * (off - 1) + 2] = ex.slot;
    }
    return n_ex;
}

int nb200_measure_fp32_peak(int device, double* tflops) {
    nb200_ctx* ctx = nullptr;
    if (!tflops) return fail(nullptr, NB200_EINVAL, "null tflops");
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || device < 0 || device >= have)
        return fail(nullptr, NB200_ECUDA, "device %d not visible (libnb200 has no CPU fallback)", device);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const int grid = prop.multiProcessorCount * 8;
    float* out = nullptr;
    CK(cudaMalloc(&out, (size_t)grid * 256 * sizeof(float)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int reps = 10;
    for (int w = 0; w < 3; ++w) nb_fma_peak_kernel<<<grid, 256>>>(out, 1.0001f);
    CK(cudaEventRecord(e0));
    for (int w = 0; w < reps; ++w) nb_fma_peak_kernel<<<grid, 256>>>(out, 1.0001f);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = (double)grid * 256.0 * NB_PEAK_ITERS * NB_PEAK_NACC * 2.0 /*lanes*/ * 2.0 /*mul+add*/ * reps;
    *tflops = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return NB200_OK;
}

int nb200_accuracy_pct(nb200_ctx* ctx, const double* forces, const double* reference, double* pct) {
    NbRange nvtx_range("nb200_accuracy_pct");
    if (!ctx || !pct || (ctx->n && !reference)) return NB200_EINVAL;
    const int D = ctx->dim;
    unsigned long long ok = 0;
    for (Shard& s : ctx->shards) {
        if (s.n_local <= 0) continue;
        if (!forces && !s.forces_valid)
            return fail(ctx, NB200_ESTATE, "accuracy of the device-resident forces needs a preceding nb200_forces call");
        CK(cudaSetDevice(s.device));
        const size_t cnt = (size_t)s.n_local * D;
        if (!s.cmp) CK(cudaMalloc(&s.cmp, 2 * cnt * sizeof(double)));
        CK(cudaMemcpyAsync(s.cmp, reference + (size_t)s.tgt_base * D, cnt * sizeof(double), cudaMemcpyHostToDevice, s.compute));
        const double* f = s.forces;
        if (forces) {
            CK(cudaMemcpyAsync(s.cmp + cnt, forces + (size_t)s.tgt_base * D, cnt * sizeof(double), cudaMemcpyHostToDevice, s.compute));
            f = s.cmp + cnt;
        }
        CK(cudaMemsetAsync(s.bounds, 0, sizeof(unsigned long long), s.compute));
        const int blocks = (int)((s.n_local + 255) / 256);
        if (D == 3) nb_accuracy_kernel<3><<<blocks, 256, 0, s.compute>>>(f, s.cmp, s.n_local, s.bounds);
        else nb_accuracy_kernel<2><<<blocks, 256, 0, s.compute>>>(f, s.cmp, s.n_local, s.bounds);
        CK(cudaGetLastError());
        ctx->launches++;
        unsigned long long h = 0;
        CK(cudaMemcpyAsync(&h, s.bounds, sizeof h, cudaMemcpyDeviceToHost, s.compute));
        CK(cudaStreamSynchronize(s.compute));
        ok += h;
    }
    *pct = ctx->n ? 100.0 * (double)ok / (double)ctx->n : 0.0;
    return NB200_OK;
}

int nb200_compare_forces(nb200_ctx* ctx, nb200_ctx* other, double* stats_out) {
    NbRange nvtx_range("nb200_compare_forces");
    if (!ctx || !other || !stats_out) return NB200_EINVAL;
    if (ctx->dim != other->dim || ctx->n != other->n || ctx->shards.size() != other->shards.size() ||
        ctx->world != other->world)
        return fail(ctx, NB200_EINVAL, "compare: the two contexts differ in dim, n or shard layout");
    for (int k = 0; k < NB200_COMPARE_STATS; ++k) stats_out[k] = 0.0;
    const int D = ctx->dim;
    for (size_t i = 0; i < ctx->shards.size(); ++i) {
        Shard& s = ctx->shards[i];
        Shard& o = other->shards[i];
        if (s.device != o.device || s.tgt_base != o.tgt_base || s.n_local != o.n_local)
            return fail(ctx, NB200_EINVAL, "compare: shard %zu lives on different devices or rows", i);
        if (s.n_local <= 0) continue;
        if (!s.forces_valid || !o.forces_valid)
            return fail(ctx, NB200_ESTATE, "compare needs a preceding nb200_forces call on both contexts");
        CK(cudaSetDevice(s.device));
        CK(cudaStreamSynchronize(o.compute));
        unsigned long long* acc = nullptr;
        CK(cudaMalloc(&acc, (2 + NB_CMP_BINS) * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(acc, 0, (2 + NB_CMP_BINS) * sizeof(unsigned long long), s.compute));
        const int blocks = (int)((s.n_local + 255) / 256);
        CK(cudaMemsetAsync(acc + 1, 0xFF, sizeof(unsigned long long), s.compute));      // argmax: minimum over the bodies at the maximum
        for (int pass = 0; pass < 2; ++pass) {
            if (D == 3) nb_compare_kernel<3><<<blocks, 256, 0, s.compute>>>(s.forces, o.forces, s.n_local, s.tgt_base, pass, acc);
            else nb_compare_kernel<2><<<blocks, 256, 0, s.compute>>>(s.forces, o.forces, s.n_local, s.tgt_base, pass, acc);
        }
        cudaError_t e = cudaGetLastError();
        unsigned long long h[2 + NB_CMP_BINS];
        if (e == cudaSuccess) e = cudaMemcpyAsync(h, acc, sizeof h, cudaMemcpyDeviceToHost, s.compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s.compute);
        cudaFree(acc);
        if (e != cudaSuccess) return fail(ctx, NB200_ECUDA, "compare kernel: %s", cudaGetErrorString(e));
        ctx->launches += 2;
        double mx;
        memcpy(&mx, &h[0], sizeof mx);
        stats_out[0] += (double)s.n_local;
        if (mx > stats_out[1]) { stats_out[1] = mx; stats_out[2] = (double)h[1]; }
        stats_out[3] += (double)h[2 + NB_CMP_BINS - 1];
        for (int k = 0; k < NB_CMP_BINS - 1; ++k) stats_out[4 + k] += (double)h[2 + k];
    }
    return NB200_OK;
}

int nb200_p2p_leaves(int device, int dim, size_t n, const void* bodies, size_t stride, size_t n_leaves,
                     const long long* leaf_offsets, const long long* leaf_bodies, const long long* nbr_offsets,
                     const long long* nbr_leaves, double G, double cutoff_r2, double eps_same, int skip_same_index, int sign,
                     double* forces_out, double* kernel_ms) {
    NbRange nvtx_range("nb200_p2p_leaves");
    nb200_ctx* ctx = nullptr;              // errors of this context-free call go to nb200_last_error(NULL)
    if (dim != 2 && dim != 3) return fail(nullptr, NB200_EINVAL, "dim must be 2 or 3, got %d", dim);
    if (sign != 1 && sign != -1) return fail(nullptr, NB200_EINVAL, "sign must be +1 (attractive, tree codes) or -1 (brute-force convention)");
    if (!(cutoff_r2 >= 0.0)) return fail(nullptr, NB200_EINVAL, "cutoff_r2 must be >= 0");
    if (n && (!bodies || !forces_out)) return fail(nullptr, NB200_EINVAL, "null bodies / forces_out");
    if (stride < (size_t)(2 * dim + 1) * sizeof(double) || stride % sizeof(double)) return fail(nullptr, NB200_EINVAL, "bad stride %zu", stride);
    if (n_leaves && (!leaf_offsets || !nbr_offsets)) return fail(nullptr, NB200_EINVAL, "null leaf / neighbour offsets");
    // the lists come from the caller's tree: check them before anything is indexed with them on the device
    const long long total = n_leaves ? leaf_offsets[n_leaves] : 0, n_nbr = n_leaves ? nbr_offsets[n_leaves] : 0;
    if (n_leaves && (leaf_offsets[0] != 0 || nbr_offsets[0] != 0 || total < 0 || n_nbr < 0))
        return fail(nullptr, NB200_EINVAL, "offset arrays must start at 0 and be non-negative");
    for (size_t l = 0; l < n_leaves; ++l)
        if (leaf_offsets[l + 1] < leaf_offsets[l] || nbr_offsets[l + 1] < nbr_offsets[l])
            return fail(nullptr, NB200_EINVAL, "offsets of leaf %zu decrease", l);
    if ((total && !leaf_bodies) || (n_nbr && !nbr_leaves)) return fail(nullptr, NB200_EINVAL, "null leaf_bodies / nbr_leaves");
    {
        std::vector<unsigned char> seen(n, 0);
        for (long long k = 0; k < total; ++k) {
            if (leaf_bodies[k] < 0 || (size_t)leaf_bodies[k] >= n) return fail(nullptr, NB200_EINVAL, "leaf_bodies[%lld] = %lld is not a body", k, leaf_bodies[k]);
            if (seen[leaf_bodies[k]]++) return fail(nullptr, NB200_EINVAL, "body %lld sits in two leaves", leaf_bodies[k]);
        }
    }
    for (long long q = 0; q < n_nbr; ++q)
        if (nbr_leaves[q] < 0 || (size_t)nbr_leaves[q] >= n_leaves) return fail(nullptr, NB200_EINVAL, "nbr_leaves[%lld] = %lld is not a leaf", q, nbr_leaves[q]);
    int have = 0;
    cudaError_t e0 = cudaGetDeviceCount(&have);
    if (e0 != cudaSuccess || have == 0)
        return fail(nullptr, NB200_ECUDA, "no CUDA device: %s (libnb200 has no CPU fallback)", e0 == cudaSuccess ? "device count is 0" : cudaGetErrorString(e0));
    if (device < 0 || device >= have) return fail(nullptr, NB200_EINVAL, "device %d not visible (%d devices)", device, have);
    for (size_t i = 0; i < n * (size_t)dim; ++i) forces_out[i] = 0.0;         // bodies in no leaf: zero, like Vector<D>()
    if (kernel_ms) *kernel_ms = 0.0;
    if (!n || !total) return NB200_OK;
    CK(cudaSetDevice(device));
    const size_t sd = stride / sizeof(double);
    double *d_aos = nullptr, *d_pos = nullptr, *d_mass = nullptr, *d_forces = nullptr;
    long long *d_lbody = nullptr, *d_loff = nullptr, *d_noff = nullptr, *d_nleaf = nullptr;
    cudaEvent_t e_start = nullptr, e_stop = nullptr;
    cudaStream_t st = nullptr;
    int rc = NB200_OK;
    auto release = [&]() {
        cudaFree(d_aos); cudaFree(d_pos); cudaFree(d_mass); cudaFree(d_forces);
        cudaFree(d_lbody); cudaFree(d_loff); cudaFree(d_noff); cudaFree(d_nleaf);
        if (e_start) cudaEventDestroy(e_start);
        if (e_stop) cudaEventDestroy(e_stop);
        if (st) cudaStreamDestroy(st);
    };
#define P2P_CK(call)                                                                                                   \
    do {                                                                                                               \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess) {                                                                                       \
            rc = fail(nullptr, NB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            release();                                                                                                 \
            return rc;                                                                                                 \
        }                                                                                                              \
    } while (0)
    P2P_CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    P2P_CK(cudaEventCreate(&e_start));
    P2P_CK(cudaEventCreate(&e_stop));
    P2P_CK(cudaMalloc(&d_aos, n * stride));
    P2P_CK(cudaMalloc(&d_pos, (size_t)dim * total * sizeof(double)));
    P2P_CK(cudaMalloc(&d_mass, (size_t)total * sizeof(double)));
    P2P_CK(cudaMalloc(&d_forces, n * (size_t)dim * sizeof(double)));
    P2P_CK(cudaMalloc(&d_lbody, (size_t)total * sizeof(long long)));
    P2P_CK(cudaMalloc(&d_loff, (n_leaves + 1) * sizeof(long long)));
    P2P_CK(cudaMalloc(&d_noff, (n_leaves + 1) * sizeof(long long)));
    P2P_CK(cudaMalloc(&d_nleaf, std::max<size_t>(1, (size_t)n_nbr) * sizeof(long long)));
    P2P_CK(cudaMemcpyAsync(d_aos, bodies, n * stride, cudaMemcpyHostToDevice, st));
    P2P_CK(cudaMemcpyAsync(d_lbody, leaf_bodies, (size_t)total * sizeof(long long), cudaMemcpyHostToDevice, st));
    P2P_CK(cudaMemcpyAsync(d_loff, leaf_offsets, (n_leaves + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    P2P_CK(cudaMemcpyAsync(d_noff, nbr_offsets, (n_leaves + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    if (n_nbr) P2P_CK(cudaMemcpyAsync(d_nleaf, nbr_leaves, (size_t)n_nbr * sizeof(long long), cudaMemcpyHostToDevice, st));
    P2P_CK(cudaMemsetAsync(d_forces, 0, n * (size_t)dim * sizeof(double), st));
    NbP2PParams Q;
    memset(&Q, 0, sizeof Q);
    Q.lpos = d_pos; Q.lmass = d_mass; Q.lbody = d_lbody; Q.leaf_off = d_loff; Q.nbr_off = d_noff; Q.nbr_leaf = d_nleaf;
    Q.forces = d_forces; Q.total = total; Q.n_leaves = (long long)n_leaves;
    Q.G = G; Q.cutoff = cutoff_r2; Q.eps_same = eps_same; Q.skip_same_index = skip_same_index; Q.sign = (double)sign;
    cudaDeviceProp prop;
    P2P_CK(cudaGetDeviceProperties(&prop, device));
    const int gb = (int)((total + 255) / 256);
    const int grid = (int)std::min<long long>((long long)n_leaves, 16LL * prop.multiProcessorCount);
    P2P_CK(cudaEventRecord(e_start, st));
    if (dim == 3) {
        nb_p2p_gather_kernel<3><<<gb, 256, 0, st>>>(d_aos, sd, d_lbody, total, d_pos, d_mass);
        nb_p2p_leaf_kernel<3><<<grid, NB_P2P_BLOCK, 0, st>>>(Q);
    } else {
        nb_p2p_gather_kernel<2><<<gb, 256, 0, st>>>(d_aos, sd, d_lbody, total, d_pos, d_mass);
        nb_p2p_leaf_kernel<2><<<grid, NB_P2P_BLOCK, 0, st>>>(Q);
    }
    P2P_CK(cudaGetLastError());
    P2P_CK(cudaEventRecord(e_stop, st));
    P2P_CK(cudaMemcpyAsync(forces_out, d_forces, n * (size_t)dim * sizeof(double), cudaMemcpyDeviceToHost, st));
    P2P_CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    P2P_CK(cudaEventElapsedTime(&ms, e_start, e_stop));
    if (kernel_ms) *kernel_ms = ms;
#undef P2P_CK
    release();
    (void)ctx;
    return NB200_OK;
}

int nb200_validation_forces(nb200_ctx* ctx, double* forces_out, long long* index_out, int cap) {
    NbRange nvtx_range("nb200_validation_forces");
    if (!ctx || cap < 0 || (cap > 0 && (!forces_out || !index_out))) return NB200_EINVAL;
    const int D = ctx->dim;
    const long long n = (long long)ctx->n;
    if (n < 3) return 0;                           // the reference divides by n / 3 (utils.h:142)
    const long long stride = n / 3;
    int count = 0;
    for (long long i = stride - 1; i < n && count < cap; i += stride) {   // (i + 1) % (n / 3) == 0
        for (Shard& s : ctx->shards) {
            if (i < s.tgt_base || i >= s.tgt_base + s.n_local) continue;
            if (!s.forces_valid)
                return fail(ctx, NB200_ESTATE, "validation forces need a preceding nb200_forces call");
            CK(cudaSetDevice(s.device));
            CK(cudaMemcpyAsync(forces_out + (size_t)count * D, s.forces + (size_t)(i - s.tgt_base) * D,
                               D * sizeof(double), cudaMemcpyDeviceToHost, s.compute));
            CK(cudaStreamSynchronize(s.compute));
            index_out[count] = i;
            ++count;
        }
    }
    return count;
}

}  // extern "C"
