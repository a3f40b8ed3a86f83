// ubench.cu -- instruction-throughput micro-benchmarks that set the FP32/FP64 pipe rooflines the
// force kernel is measured against (FFMA2 throughput was "still to be measured").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu
// Each kernel runs ONE wave of persistent CTAs; SM clock = delta clock64 / delta globaltimer
// measured inside the kernel, so the per-clock rates do not depend on launch overhead.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 8192
#define NACC 8

__device__ __forceinline__ float rcpf(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

enum { FFMA_S, FFMA2_REUSE, FFMA2_3DIST, FFMA2_SQ, FADD2_BC, FMUL2_SQ, MIX, MIX_NOCUT, DFMA_, DFMA_3DIST, MUFU_, FSEL_, MIX_MIN, MIX_RCP4, MIX_SACC, MIX_SACC1, MIX_SACC2 };

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, unsigned long long* stamp, float a, float b) {
    float2 acc[NACC], x[NACC], y[NACC];
    double dacc[NACC], dx[NACC], dy[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
        x[i] = make_float2(a + i * 1e-6f, a - i * 1e-6f);
        y[i] = make_float2(b + i * 1e-9f, b - i * 1e-9f);
        dacc[i] = threadIdx.x + i; dx[i] = a + i * 1e-9; dy[i] = b + i * 1e-12;
    }
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
    float rmin = 1e30f;
    __syncthreads();
    unsigned long long g0 = gtime(); long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == FFMA_S) { acc[i].x = fmaf(acc[i].x, a, b); acc[i].y = fmaf(acc[i].y, a, b); }
            else if (MODE == FFMA2_REUSE) acc[i] = __ffma2_rn(acc[i], a2, b2);
            else if (MODE == FFMA2_3DIST) acc[i] = __ffma2_rn(x[i], y[i], acc[i]);
            else if (MODE == FFMA2_SQ) acc[i] = __ffma2_rn(x[i], x[i], acc[i]);
            else if (MODE == FADD2_BC) acc[i] = __fadd2_rn(acc[i], make_float2(a, a));
            else if (MODE == FMUL2_SQ) acc[i] = __fmul2_rn(acc[i], acc[i]);
            else if (MODE == DFMA_) dacc[i] = fma(dacc[i], (double)a, (double)b);
            else if (MODE == DFMA_3DIST) dacc[i] = fma(dx[i], dy[i], dacc[i]);
            else if (MODE == MUFU_) acc[i].x = rcpf(acc[i].x);
            else if (MODE == FSEL_) { acc[i].x = acc[i].x >= a ? acc[i].x : b; acc[i].y = acc[i].y >= b ? acc[i].y : a; a += 1e-9f; }
            else if (MODE == MIX || MODE == MIX_NOCUT || MODE == MIX_MIN || MODE == MIX_RCP4 || MODE == MIX_SACC || MODE == MIX_SACC1 || MODE == MIX_SACC2) {
                // one packed pair-chain of the force kernel per i: 3 FADD2, FMUL2, 2 FFMA2, [2 FSETP+2 FSEL], 2 MUFU,
                // 2 FMUL2, 3 FFMA2  (11 FMA-pipe packed ops)
                const float2 s0 = make_float2(a + it, b + it);
                const float2 d0 = __fadd2_rn(s0, x[i]), d1 = __fadd2_rn(s0, y[i]), d2 = __fadd2_rn(a2, x[i]);
                float2 r2 = __fmul2_rn(d0, d0); r2 = __ffma2_rn(d1, d1, r2); r2 = __ffma2_rn(d2, d2, r2);
                if (MODE == MIX) { r2.x = r2.x >= b ? r2.x : 1e38f; r2.y = r2.y >= b ? r2.y : 1e38f; }
                if (MODE == MIX_MIN) rmin = fminf(rmin, fminf(r2.x, r2.y));
                float2 s;
                if (MODE == MIX_RCP4) { const float2 r4 = __fmul2_rn(r2, r2); s = make_float2(rcpf(r4.x), rcpf(r4.y)); s = __fmul2_rn(s, b2); }
                else { float2 inv = make_float2(rcpf(r2.x), rcpf(r2.y)); s = __fmul2_rn(inv, inv); s = __fmul2_rn(s, b2); }
                // accumulate: packed FFMA2 has three distinct 64-bit operands (register-fetch bound);
                // MIX_SACC* replace 3/1/2 of them by two scalar FFMA each
                constexpr int NS = MODE == MIX_SACC ? 3 : MODE == MIX_SACC2 ? 2 : MODE == MIX_SACC1 ? 1 : 0;
                if (NS >= 1) { acc[i].x = fmaf(s.x, d0.x, acc[i].x); acc[i].y = fmaf(s.y, d0.y, acc[i].y); }
                else acc[i] = __ffma2_rn(s, d0, acc[i]);
                if (NS >= 2) { acc[(i + 1) % NACC].x = fmaf(s.x, d1.x, acc[(i + 1) % NACC].x); acc[(i + 1) % NACC].y = fmaf(s.y, d1.y, acc[(i + 1) % NACC].y); }
                else acc[(i + 1) % NACC] = __ffma2_rn(s, d1, acc[(i + 1) % NACC]);
                if (NS >= 3) { acc[(i + 2) % NACC].x = fmaf(s.x, d2.x, acc[(i + 2) % NACC].x); acc[(i + 2) % NACC].y = fmaf(s.y, d2.y, acc[(i + 2) % NACC].y); }
                else acc[(i + 2) % NACC] = __ffma2_rn(s, d2, acc[(i + 2) % NACC]);
            }
        }
    }
    long long t1 = clock64(); unsigned long long g1 = gtime();
    float s = 0; double ds = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { s += acc[i].x + acc[i].y; ds += dacc[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)ds + a + rmin;
    if (threadIdx.x == 0) { stamp[2 * blockIdx.x] = (unsigned long long)(t1 - t0); stamp[2 * blockIdx.x + 1] = g1 - g0; }
}

template <int MODE> void run(const char* name, double lane_ops_per_inner, int ctas_per_sm) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * ctas_per_sm;
    float* out; unsigned long long* st; cudaMalloc(&out, grid * 256 * sizeof(float)); cudaMalloc(&st, grid * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) k<MODE><<<grid, 256>>>(out, st, 1.0001f, 1e-7f);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int w = 0; w < reps; ++w) k<MODE><<<grid, 256>>>(out, st, 1.0001f, 1e-7f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long* h = new unsigned long long[2 * grid];
    cudaMemcpy(h, st, grid * 16, cudaMemcpyDeviceToHost);
    double cyc = 0, ns = 0;
    for (int i = 0; i < grid; ++i) { cyc += h[2 * i]; ns += h[2 * i + 1]; }
    cyc /= grid; ns /= grid;
    const double mhz = cyc / ns * 1e3;
    const double ops_cta = 256.0 * ITERS * NACC * lane_ops_per_inner;      // lane-ops per CTA
    const double per_clk_sm = ops_cta * ctas_per_sm / cyc;
    printf("%-34s ctas/SM=%d  %7.3f ms  %7.2f T lane-ops/s  %7.2f lane-ops/clk/SM  SMclk %4.0f MHz  %s\n", name, ctas_per_sm,
           ms / reps, ops_cta * grid / (ms / reps * 1e-3) / 1e12, per_clk_sm, mhz, cudaGetErrorString(cudaGetLastError()));
    delete[] h; cudaFree(out); cudaFree(st);
}

int main() {
    for (int c : {2, 4}) {
        run<FFMA_S>("FFMA scalar x2", 2, c);
        run<FFMA2_REUSE>("FFMA2 acc=acc*a+b (reuse)", 2, c);
        run<FFMA2_3DIST>("FFMA2 acc=x*y+acc (3 distinct)", 2, c);
        run<FFMA2_SQ>("FFMA2 acc=x*x+acc", 2, c);
        run<FADD2_BC>("FADD2 acc+=bcast", 2, c);
        run<FMUL2_SQ>("FMUL2 acc*=acc", 2, c);
        run<MIX>("pair chain (11 packed+2MUFU+4ALU)", 22, c);
        run<MIX_NOCUT>("pair chain no cut-off", 22, c);
        run<MIX_MIN>("pair chain + FMNMX3 min", 22, c);
        run<MIX_RCP4>("pair chain rcp(r2*r2), no cut", 22, c);
        run<MIX_SACC1>("pair chain, 1 of 3 acc scalar", 22, c);
        run<MIX_SACC2>("pair chain, 2 of 3 acc scalar", 22, c);
        run<MIX_SACC>("pair chain, all acc scalar FFMA", 22, c);
        run<DFMA_>("DFMA reuse", 1, c);
        run<DFMA_3DIST>("DFMA 3 distinct", 1, c);
        run<MUFU_>("MUFU.RCP", 1, c);
        run<FSEL_>("FSETP+FSEL x2", 2, c);
    }
    return 0;
}
