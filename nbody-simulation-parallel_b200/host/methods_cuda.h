// methods_cuda.h -- the B200 brute-force path behind the reference's methods.h interface.
//
// Drop-in for /root/reference/nbody-sim-new: same types (Body<D> body.h:7-19, Vector<D>
// vector.h:9-12), same calling convention and return type as the five CPU brute-force entry
// points (methods.h:29-43), same integrator semantics as update_body_velocities /
// update_body_positions (methods.h:85-91, methods.cpp:426-450), errors as C++ exceptions so that
// safely_execute (utils.h:87-104) logs them and skips the CSV row.  Host code only: all device
// work happens behind the extern "C" layer of libnb200.so (include/nb200.h).
//
//   forces = brute_force_cuda_n_body<D>(bodies);            // one force evaluation, like its peers
//   brute_force_cuda_simulate<D>(bodies, dt, steps);        // force + v += F/m dt + x += v dt, fused
//   brute_force_cuda_warmup<D>(n);                          // optional: CUDA/NCCL init outside the timed call
//
// Run-time selection (environment, read once per session):
//   NB200_PRECISION = 64 (default, the reference's precision) | 32 (packed-FP32 pair math)
//   NB200_GPUS      = 1 (default) .. 8   targets sharded over the GPUs of the box
// G and the pair cut-off are the reference's own constants (utils.h:21; methods.cpp:24).
#ifndef METHODS_CUDA_H
#define METHODS_CUDA_H

#include <cstddef>
#include <vector>

#include "body.h"
#include "vector.h"

template <int D>
std::vector<Vector<D>> brute_force_cuda_n_body(const std::vector<Body<D>>& bodies);

template <int D>
void brute_force_cuda_simulate(std::vector<Body<D>>& bodies, double dt, int steps);

// Creates (and caches) the device context for n bodies so that the first timed call does not pay
// for CUDA context creation, cudaMalloc or ncclCommInitAll.  Safe to skip.
template <int D>
void brute_force_cuda_warmup(std::size_t n);

// Device time (ms, CUDA events) of the kernels of the last call -- the orphan main_cuda.cu's
// timing convention (main_cuda.cu:123-137), reported beside safely_execute's wall clock.
double brute_force_cuda_last_kernel_ms();

// Releases the cached device context (also done at process exit).
void brute_force_cuda_release();

#endif  // METHODS_CUDA_H
