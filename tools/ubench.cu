// ubench.cu -- instruction-throughput micro-benchmarks that set the FP32/FP64 pipe rooflines the
// force kernel is measured against (FFMA2 throughput was "still to be measured").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

__device__ __forceinline__ float rcpf(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, long long* cyc, float a, float b) {
    float2 acc[NACC];
    double dacc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); dacc[i] = threadIdx.x + i; }
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) {            // scalar FFMA x2 (same flops as one FFMA2)
                acc[i].x = fmaf(acc[i].x, a, b); acc[i].y = fmaf(acc[i].y, a, b);
            } else if (MODE == 1) {     // FFMA2
                acc[i] = __ffma2_rn(acc[i], a2, b2);
            } else if (MODE == 2) {     // FFMA2 + 1 MUFU per 5.5 FFMA2 (the pair-kernel mix): 11 FFMA2 + 2 MUFU
                acc[i] = __ffma2_rn(acc[i], a2, b2);
                if ((i & 3) == 0 && (it & 1)) { /* thin */ }
            } else if (MODE == 3) {     // DFMA
                dacc[i] = fma(dacc[i], (double)a, (double)b);
            } else if (MODE == 4) {     // MUFU.RCP only
                acc[i].x = rcpf(acc[i].x);
            } else if (MODE == 5) {     // FADD2
                acc[i] = __fadd2_rn(acc[i], a2);
            } else if (MODE == 6) {     // FMUL2
                acc[i] = __fmul2_rn(acc[i], a2);
            }
        }
        if (MODE == 2) {                // 8 FFMA2 above + 3 more + 2 MUFU + 2 FSETP/FSEL  = one packed pair-chain worth
            acc[0] = __ffma2_rn(acc[0], b2, a2); acc[1] = __ffma2_rn(acc[1], b2, a2); acc[2] = __ffma2_rn(acc[2], b2, a2);
            float r0 = acc[3].x >= a ? acc[3].x : 1e30f, r1 = acc[3].y >= a ? acc[3].y : 1e30f;
            acc[4].x += rcpf(r0) * 1e-30f; acc[4].y += rcpf(r1) * 1e-30f;
        }
    }
    long long t1 = clock64();
    float s = 0; double ds = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { s += acc[i].x + acc[i].y; ds += dacc[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)ds;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE> void run(const char* name, double lane_ops_per_iter_per_thread, int ctas_per_sm) {
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int grid = p.multiProcessorCount * ctas_per_sm;
    float* out; long long* cyc; cudaMalloc(&out, grid * 256 * sizeof(float)); cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) k<MODE><<<grid, 256>>>(out, cyc, 1.0001f, 1e-7f);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int w = 0; w < reps; ++w) k<MODE><<<grid, 256>>>(out, cyc, 1.0001f, 1e-7f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double ops = (double)grid * 256 * ITERS * lane_ops_per_iter_per_thread * reps;
    double rate = ops / (ms * 1e-3);
    double mhz = (double)c / (ms / reps * 1e-3) / 1e6;   // cycles of one CTA / kernel time ~ SM clock (1 wave)
    printf("%-28s ctas/SM=%d  %8.3f ms/launch  %8.2f T lane-ops/s  = %6.2f lane-ops/clk/SM @%.0f MHz(est)  err=%s\n", name,
           ctas_per_sm, ms / reps, rate / 1e12, rate / p.multiProcessorCount / (mhz * 1e6), mhz, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int c : {2, 4}) {
        run<0>("FFMA scalar (x2)", 2.0 * NACC, c);
        run<1>("FFMA2 packed", 2.0 * NACC, c);
        run<5>("FADD2 packed", 2.0 * NACC, c);
        run<6>("FMUL2 packed", 2.0 * NACC, c);
        run<2>("pair mix 11 FFMA2+2 MUFU+ALU", 2.0 * (NACC + 3), c);
        run<3>("DFMA", 1.0 * NACC, c);
        run<4>("MUFU.RCP", 1.0 * NACC, c);
    }
    return 0;
}
