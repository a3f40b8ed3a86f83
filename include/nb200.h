/*
 * nb200.h -- C ABI of libnb200.so: the B200-native (sm_100a) brute-force N-body path.
 *
 * This is the drop-in boundary for ONE hot path of mathaiml5/NBody-simulation-parallel
 * (paths below are relative to /root/reference/nbody-sim-new):
 *
 *   nb200_forces        replaces  brute_force_seq_n_body<D>       methods.h:30-31  methods.cpp:7-42
 *                                 brute_force_omp_n_body_1<D>     methods.h:33-34  methods.cpp:45-95
 *                                 brute_force_omp_n_body_2<D>     methods.h:36-37  methods.cpp:98-136
 *                                 brute_force_parlay_n_body_1<D>  methods.h:39-40  methods.cpp:139-186
 *                                 brute_force_parlay_n_body_2<D>  methods.h:42-43  methods.cpp:189-224
 *   nb200_step          replaces  the loop {force; update_body_velocities<D>; update_body_positions<D>}
 *                                 methods.h:85-91  methods.cpp:426-450   (semi-implicit Euler)
 *   nb200_upload_aos /  take and return the reference's own AoS std::vector<Body<D>> bytes
 *   nb200_download_aos            body.h:7-19  (stride 40 B for D=2, 56 B for D=3), vector.h:9-12
 *   nb200_accuracy_pct  replaces  compute_accuracy_omp<D>         utils.h:170-219
 *
 * Semantics kept exactly as the reference codes them: F_i = -G m_i sum_j m_j (p_j - p_i) / r^4
 * (methods.cpp:21-37), pairs with r^2 < cutoff DROPPED (methods.cpp:24; the reference hard-codes
 * 1e-10), forces (not accelerations) returned in body order as D doubles per body, G and the
 * cut-off passed at run time (the reference: utils.h:21, G = 4.471e-21).
 *
 * Plain pointers and sizes only; no exceptions cross this boundary; every entry point returns
 * 0 on success or a negative NB200_E* code, with text from nb200_last_error().  There is no CPU
 * fallback: without a usable sm_100 device every compute entry point fails with NB200_ECUDA.
 *
 * The C++ adapter with the reference's template signatures is
 * nbody-simulation-parallel_b200/host/methods_cuda.h; INTEGRATION.md shows the binding.
 */
#ifndef NB200_H
#define NB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nb200_ctx nb200_ctx;

/* Precision of the PAIR arithmetic.  Error metric everywhere: per body i, ||F_i - F_ref,i||_2 / ||F_ref,i||_2.
 *
 * NB200_FP64  double throughout: <= 1e-12 against the reference's brute-force methods on identical inputs.
 * NB200_FP32  packed-FP32 pair math on the inputs QUANTISED to 24 bits (positions x * 2^-p, masses m * 2^-k, exact
 *             power-of-two scales); accumulation, state and integration stay FP64.  Against the reference evaluated
 *             on the same quantised inputs:  <= max(1e-5, 6e-7 * kappa_i),  kappa_i = sum_j |f_ij| / |sum_j f_ij|
 *             (the body's own summation condition number): the flat 1e-5 for every body with kappa_i <= 16.7
 *             (measured on all 2^20 bodies of the headline config: 4 bodies above 1e-5, maximum 1.6e-5).
 *             Against the reference on the UNROUNDED double inputs the position quantisation itself moves each
 *             near-neighbour term by ~4 * 2^-24 * |x| / r: at N = 2^20 in the unit cube 35 % of the bodies differ by
 *             more than 1e-5 (maximum 1.8e-3; nb200_compare_forces reports the histogram).  Where that matters use
 *             NB200_FP64, or option "fp32_positions" = 48 (below): the same FP32 pair arithmetic on positions kept as
 *             float PAIRS (hi + lo), differences taken as (hi_j - hi_i) + (lo_j - lo_i).  Then the band above holds
 *             against the reference on the UNROUNDED inputs (measured on the same 2^20 bodies: 4 above 1e-5, maximum
 *             2.3e-5) at 74 % of the FP32 throughput (2900 vs 3904 G interactions/s; FP64: 1706).  One GPU, pair-
 *             symmetric pass.  The reference's own -a 1 criterion (1 % per component, utils.h:170-219) is met by all. */
#define NB200_FP64 64
#define NB200_FP32 32

#define NB200_OK 0
#define NB200_EINVAL (-1) /* bad argument */
#define NB200_ECUDA (-2)  /* CUDA runtime error / no device */
#define NB200_ENCCL (-3)  /* NCCL error or NCCL not loadable */
#define NB200_ESTATE (-4) /* call sequence error (e.g. step before upload) */
#define NB200_ENOMEM (-5)

#define NB200_UNIQUE_ID_BYTES 128
#define NB200_IPC_BYTES 256

/* ---- lifetime --------------------------------------------------------------------------- */

/* Single-process context over `ngpus` devices (0..ngpus-1, or the list in env NB200_DEVICES):
 * targets are sharded by contiguous index range, one NCCL communicator per device
 * (ncclCommInitAll) when ngpus > 1.  dim = 2 or 3; n = number of bodies. */
int nb200_create(nb200_ctx** out, int dim, size_t n, int precision, int ngpus);

/* One-process-per-GPU context (torchrun style): this process owns shard `rank` of `world` on
 * CUDA device `device`.  unique_id = NB200_UNIQUE_ID_BYTES bytes from nb200_get_unique_id() on
 * rank 0, broadcast by the caller (any transport); may be NULL when world == 1.  NULL with
 * world > 1 builds a DETACHED shard without a communicator (test hook): nb200_forces works,
 * nb200_step is limited to nsteps == 1 and the caller re-uploads before the next step. */
int nb200_create_rank(nb200_ctx** out, int dim, size_t n, int precision, int device, int rank,
                      int world, const void* unique_id);
int nb200_get_unique_id(void* unique_id_out);

/* Fused NVLink exchange for rank contexts (the default whenever it is attached): instead of an
 * all-gather after each step, the integrator epilogue stores the new source rows straight into
 * every peer's next-step buffer over NVLink peer memory, and a per-pair flag publishes the step.
 * Each rank exports one NB200_IPC_BYTES blob (CUDA IPC handles of its two source buffers and its
 * flag array); the caller all-gathers the blobs in rank order (any transport) and every rank
 * attaches all `world` of them.  Single-process contexts (nb200_create) wire this up themselves
 * through CUDA peer access.  Needs world <= 8 and peer access between all GPUs; without it the
 * context keeps using its NCCL communicator. */
int nb200_ipc_export(nb200_ctx* ctx, void* blob_out);
int nb200_ipc_attach(nb200_ctx* ctx, const void* blobs, int count);

void nb200_destroy(nb200_ctx* ctx);

/* ---- data ------------------------------------------------------------------------------- */

/* bodies = n records of the reference's Body<D>: position[D], velocity[D], mass (doubles);
 * stride in bytes (40 for D=2, 56 for D=3, or larger).  Every rank passes ALL n bodies. */
int nb200_upload_aos(nb200_ctx* ctx, const void* bodies, size_t stride);

/* Seeded synthetic bodies generated ON the device instead of uploaded (same state afterwards as
 * nb200_upload_aos of the same bodies with the packed 40/56-byte stride; read them back with
 * nb200_download_aos).  The reference's generate_random_bodies<D> (utils.h:107-135) is unseeded;
 * kind 0 reproduces its ranges (pos U[1,1e7], vel U[-10,10], mass U[1,1e8], utils.h:113-115), kind 1
 * is the uniform unit cube/square (vel U[-0.1,0.1], mass U[0.5,1.5]/(G n)), kind 2 the 3D Plummer
 * sphere (equal masses 1/(G n)).  Bodies depend only on (kind, seed, n, G, body index) --
 * Philox-4x32-10 -- so every rank of a sharded run generates the same set. */
#define NB200_GEN_REFERENCE_RANGE 0
#define NB200_GEN_UNIFORM 1
#define NB200_GEN_PLUMMER 2
int nb200_generate(nb200_ctx* ctx, int kind, unsigned long long seed, double G);

/* Writes position and velocity of the bodies this context owns back into an AoS array of ALL n
 * bodies: every body for nb200_create contexts, rows [lo,hi) of nb200_shard_range for
 * nb200_create_rank contexts.  Same stride as the upload.  With packed records (stride 40 / 56)
 * whole rows are copied, i.e. the mass field is rewritten with the uploaded mass (a step never
 * changes it); with padded records only position and velocity are written. */
int nb200_download_aos(nb200_ctx* ctx, void* bodies, size_t stride);

/* Owned target range [lo, hi) in body indices. */
int nb200_shard_range(const nb200_ctx* ctx, size_t* lo, size_t* hi);

/* ---- the hot path ----------------------------------------------------------------------- */

/* cutoff_r2: pairs with r^2 < cutoff_r2 are dropped (methods.cpp:24 hard-codes 1e-10).  Values below
 * 1e-20 are clamped to 1e-20: Vector<D>::normalized() already zeroes every pair with r < 1e-10
 * (vector.h:95), so the reference cannot express a smaller cut-off either.
 *
 * One force evaluation at the uploaded positions.  forces_out = n*D doubles (Vector<D> layout,
 * body order); rank contexts fill rows [lo,hi) only.  Does not change the state. */
int nb200_forces(nb200_ctx* ctx, double G, double cutoff_r2, double* forces_out);

/* nsteps of: F = brute force; v += (F/m) dt; x += v dt  -- entirely on the device(s), integrator
 * fused into the force kernel's epilogue, one position all-gather per step when sharded. */
int nb200_step(nb200_ctx* ctx, double G, double cutoff_r2, double dt, int nsteps);

/* Kinetic and potential energy of the reference's own law (U = sum_{i<j} G m_i m_j / (2 r^2),
 * pairs under the cut-off excluded).  Rank contexts return their shard's partial sums. */
int nb200_energy(nb200_ctx* ctx, double G, double cutoff_r2, double* kinetic, double* potential);

/* Test aid (pure host logic, no device needed): the work list of the pair-symmetric pass for
 * `rank` of `world` over n bodies, 4 ints per row: {i-tile, first source tile, end source tile,
 * flags (bit 0: pair-symmetric row; clear: ordered pairs)}.  i-tiles are 1024 targets = 4 source
 * tiles of 256, counted from the rank's first tile.  Writes at most `cap` rows, returns the row
 * count (or a negative NB200_E* code). */
int nb200_debug_sym_rows(size_t n, int world, int rank, int* rows_out, int cap);

/* Test aid (pure host logic): the reaction-sum exchange of `rank` in the cross-rank pair-symmetric
 * pass, 3 ints per partner offset 1..floor(world/2): {rank it pushes to, rank it receives from, slot
 * index (in the receiver's memory for the push, in its own for the receive)}.  Returns floor(world/2). */
int nb200_debug_sym_exchange(int world, int rank, int* out, int cap);

/* Measurement aid: the FP32 FMA-pipe throughput this device sustains on independent packed
 * FFMA2 chains (TFLOP/s, 2 flops per lane-op), timed with CUDA events.  It is the denominator of
 * the FP32 roofline fraction bench.py reports next to the nominal SMs x 128 x 2 x clock figure. */
int nb200_measure_fp32_peak(int device, double* tflops);

/* compute_accuracy_omp<D> (utils.h:170-219), evaluated on the device: percentage of bodies whose
 * every force component is within 1 % of the reference component (|reference| < 1e-20: |force|
 * <= 1e-9 instead).  `reference` = n*D doubles on the host.  `forces` = n*D doubles on the host, or
 * NULL to compare the forces of the last nb200_forces call that are still on the device (no
 * download).  Rank contexts count their own rows only: the per-rank values add up to the total. */
int nb200_accuracy_pct(nb200_ctx* ctx, const double* forces, const double* reference, double* pct);

/* Full-population parity metric on the device: per-body norm-wise relative difference
 * ||F_a - F_b|| / ||F_b|| between the forces of the last nb200_forces call of two contexts over the
 * same n bodies (same dim, same shard layout and devices), e.g. an NB200_FP32 context against an
 * NB200_FP64 one -- the measurement compute_accuracy_omp's 1 % criterion (utils.h:170-219) is too
 * coarse for.  stats_out = NB200_COMPARE_STATS doubles:
 *   [0] bodies compared  [1] maximum difference  [2] body index of the maximum  [3] non-finite count
 *   [4 + k], k = 0..17: bodies with difference in [10^(k-17), 10^(k-16)) (k = 0: everything below 1e-16)
 * Rank contexts compare their own rows; counts add up over ranks, the maximum is the max. */
#define NB200_COMPARE_STATS 22
int nb200_compare_forces(nb200_ctx* a, nb200_ctx* b, double* stats_out);

/* print_validation_forces<D> (utils.h:138-151) without downloading all forces: the forces of the
 * bodies the reference prints (0-based i with (i+1) % (n/3) == 0) from the last nb200_forces
 * call.  forces_out = cap*D doubles, index_out = cap body indices; returns the number of bodies
 * written (rank contexts: the ones they own), 0 when n < 3, or a negative NB200_E* code. */
int nb200_validation_forces(nb200_ctx* ctx, double* forces_out, long long* index_out, int cap);

/* ---- next row of the suite: the P2P (leaf) step of its tree codes --------------------------------- */

/* Direct sums over LEAF LISTS, the particle-to-particle step of the reference's Barnes-Hut / BVH / FMM methods:
 *   BVH<D>::calculate_force, leaf branch            bvh.cpp:149-177   (eps_same = 1e-9, cutoff_r2 = 1e-9)
 *   FMM<D>::calculate_accurate_force, leaf branch   fmm.cpp:622-637   (skip_same_index = 1, cutoff_r2 = 1e-10)
 *   fmm_parlay.cpp:918-1023                                           (same law and guards)
 * The pair law is the brute-force methods' (|F| = G m_i m_j / r^3 along d = p_j - p_i), the sign the tree codes'
 * (sign = +1: force += diff.normalized() * mag, attractive; -1 gives the brute-force convention).
 *   bodies        n records of Body<D> (body.h:7-19), `stride` bytes apart
 *   leaf_offsets  [n_leaves + 1] into leaf_bodies; leaf_bodies = body indices, leaf after leaf (a body sits in at most
 *                 one leaf)
 *   nbr_offsets   [n_leaves + 1] into nbr_leaves; nbr_leaves = the SOURCE leaves of each target leaf (the leaf itself
 *                 included when its own bodies interact, as in the reference's leaf loop)
 *   eps_same      >= 0: a pair is skipped when every |d_k| <= eps_same (bvh.cpp:156-163); < 0: test off
 *   cutoff_r2     pairs with r^2 < cutoff_r2 are skipped
 *   skip_same_index  1: a body never interacts with itself even at cutoff_r2 = 0 (fmm.cpp:624)
 *   forces_out    n * dim doubles, Vector<D> layout, body order; bodies in no leaf get zero
 *   kernel_ms     optional: device time of the gather + P2P kernels (CUDA events)
 * FP64 throughout (<= 1e-12 against the reference's own leaf loop).  Context-free: errors via nb200_last_error(NULL). */
int nb200_p2p_leaves(int device, int dim, size_t n, const void* bodies, size_t stride, size_t n_leaves,
                     const long long* leaf_offsets, const long long* leaf_bodies, const long long* nbr_offsets,
                     const long long* nbr_leaves, double G, double cutoff_r2, double eps_same, int skip_same_index,
                     int sign, double* forces_out, double* kernel_ms);

/* ---- introspection / measurement --------------------------------------------------------- */

/* Device time (CUDA events on the launching stream, max over this context's devices) of the
 * kernels of the last nb200_forces / nb200_step call, in milliseconds. */
int nb200_last_elapsed_ms(const nb200_ctx* ctx, double* ms);

/* Kernel launches issued by this context since creation (our own kernels only). */
long long nb200_launch_count(const nb200_ctx* ctx);

/* Options (all optional; the defaults are the measured choices):
 *   "deterministic" 1 = bit-reproducible sums: the ordered pass writes its unit partial sums to per-segment slots and adds
 *                 them in segment order instead of meeting in FP64 atomics.  Identical bits from run to run AND between
 *                 1 and N GPUs / shards (every target's sum is the same expression on any shard count).  Costs the
 *                 pair-symmetric pass (2.8 instead of 3.7 T interactions/s at N = 2^20).  Default 0.
 *   "equal_mass"  0 = never use the equal-mass flavour of the pair kernels (default: used when every body has the same
 *                 mass); set before the upload
 *   "symmetric"   1/0 = pair-symmetric pass on/off (default: on from 12288 bodies): every unordered pair is evaluated once
 *                 and feeds both bodies, like the j > i loop of brute_force_seq_n_body (methods.cpp:18-39); across ranks the
 *                 reaction sums are pushed to their owner over NVLink (needs the fused exchange)
 *   "detect"      1/0 = close-pair pre-pass (hash grid) on/off (default: on from 49152 bodies); without it the
 *                 pair-symmetric pass applies the exact cut-off to every pair, the ordered pass tracks the minimum r^2
 *   "sym_algo"    how the pair-symmetric kernel sums the reactions on the streamed sources: 0 shared-memory transpose,
 *                 1 register rotation through the warp, 2 (FP32 default) rotation with decoupled hand-over
 *   "sym_ti", "sym_block"  register-block shape of the pair-symmetric kernel: targets per thread x threads per CTA.
 *                 FP32: 8 x 128 (default), 4 x 256, 4 x 128 (default up to 24576 bodies per shard); FP64: 8 x 128 (default),
 *                 4 x 256, 4 x 128, 2 x 256, 2 x 128 (the 8 x 128 and 4 x 128 FP64 shapes exist for the rotation, sym_algo 1)
 *   "seg_tiles" / "seg_sub"  source tiles / sub-tiles (128 sources FP32, 64 FP64) per work unit
 *   "variant"     index into the compiled (targets/thread, j-split, block) table of the ordered pass, -1 = auto
 *   "grid_mult"   ordered pass: persistent CTAs = grid_mult * (SMs * occupancy) / 16  (16 = exactly resident)
 *   "exchange"    1 = fused NVLink peer stores from the epilogue (default when attached), 0 = ncclAllGather
 *   "overlap"     ordered pass on several GPUs: 1 = own source rows first (the peer handshake is taken when a CTA first needs
 *                 remote rows; NCCL: local pass | all-gather | remote pass), 0 = one pass after the handshake, -1 = auto
 *   "shard_upload" 1 (default with the fused exchange) = nb200_upload_aos copies only the rows a shard owns and stores
 *                 their packed source rows into every peer's buffers (the call is then collective over the ranks), 0 = every
 *                 rank uploads and packs all n bodies
 *   "pdl"         0 = no programmatic dependent launch between the pair-symmetric pass and its finish kernel (default on)
 *   "spin_timeout_ms"  bound of every device-side wait on a peer's flag (default 30000): when it expires the call returns
 *                 NB200_ESTATE naming the peer and what was awaited, and the context refuses further work
 *   "fp32_positions"  24 (default) or 48, NB200_FP32 contexts on one GPU, set BEFORE the upload: 48 keeps every scaled
 *                 coordinate as two floats (the second row is written by the pack kernel and by the integrator every
 *                 step) so that the 24-bit quantisation of the positions no longer moves the near field; the close-pair
 *                 pre-pass runs on the hi parts with cells widened by their largest difference error; excludes
 *                 "deterministic"
 *   "trace"       1 = append a CUDA-event timeline of shard 0 to nb200_plan() after nb200_step.  Independently of it every
 *                 entry point that touches the device opens an NVTX range of its own name (visible in Nsight Systems;
 *                 a no-op without a tool attached)
 *   "sym_itile"   256 = one-source-tile i-tiles (4 x 64 threads, transpose flavour): a small-N experiment kept for measurements
 *   "debug_fake_peer"  test hook for the time-out path on one GPU (a detached shard waits for a peer that does not exist)
 * All ranks of a sharded run must set the same options.  Returns NB200_EINVAL for an unknown key. */
int nb200_set_option(nb200_ctx* ctx, const char* key, long value);

/* Human-readable one-line description of the launch plan of the last call (for logs). */
const char* nb200_plan(const nb200_ctx* ctx);

const char* nb200_last_error(const nb200_ctx* ctx); /* ctx may be NULL: last create() error */
const char* nb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NB200_H */
