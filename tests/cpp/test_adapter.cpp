// test_adapter.cpp -- exercises the C++ adapter (methods_cuda.h) the way run_benchmark<D> does
// (main.cpp:136-140) and checks it against the oracle (liboracle.so; tests may use the checker).
// Usage: test_adapter <dim> <n> ; prints "ADAPTER_OK ..." and exits 0 on success.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <stdexcept>
#include <vector>

#include "methods_cuda.h"

extern "C" int oracle_forces_omp2(int D, size_t n, const double* bodies, double G, double cutoff, double* forces);
extern "C" int oracle_simulate(int D, size_t n, double* bodies, double G, double cutoff, double dt, int nsteps,
                               int variant);
extern "C" int oracle_condition(int D, size_t n, const double* bodies, double G, double cutoff, double* kappa);

template <int D>
int run(size_t n) {
    // the reference generator's ranges (utils.h:113-115), seeded
    std::mt19937_64 gen(1234);
    std::uniform_real_distribution<double> pos(1.0, 1.0e7), vel(-10.0, 10.0), mass(1.0, 1.0e8);
    std::vector<Body<D>> bodies(n);
    for (auto& b : bodies) {
        for (int d = 0; d < D; ++d) { b.position[d] = pos(gen); b.velocity[d] = vel(gen); }
        b.mass = mass(gen);
    }
    const char* prec = std::getenv("NB200_PRECISION");
    const bool fp32 = prec && std::atoi(prec) == 32;
    const bool fp48 = prec && std::atoi(prec) == 48;   // FP32 arithmetic on 48-bit positions: same band, UNROUNDED inputs
    const bool band = fp32 || fp48;
    if (fp32) {   // FP32 mode is specified against the oracle fed float-rounded inputs (tests/test_gpu_parity.py)
        for (auto& b : bodies) {
            for (int d = 0; d < D; ++d) b.position[d] = (double)(float)b.position[d];
            b.mass = (double)(float)b.mass;
        }
    }
    brute_force_cuda_warmup<D>(n);
    std::vector<Vector<D>> f = brute_force_cuda_n_body<D>(bodies);
    std::vector<double> ref(n * D);
    oracle_forces_omp2(D, n, reinterpret_cast<const double*>(bodies.data()), 4.471e-21, 1e-10, ref.data());
    // the same criterion as tests/test_gpu_parity.py and include/nb200.h: per-body norm-wise relative error,
    // FP64 <= 1e-12; FP32 <= max(1e-5, 6e-7 * kappa_i) with the body's summation condition number kappa_i
    std::vector<double> kappa(n, 1.0);
    if (band) oracle_condition(D, n, reinterpret_cast<const double*>(bodies.data()), 4.471e-21, 1e-10, kappa.data());
    double worst = 0.0;          // worst error relative to the body's bound
    for (size_t i = 0; i < n; ++i) {
        double num = 0.0, den = 0.0;
        for (int d = 0; d < D; ++d) {
            const double e = f[i][d] - ref[i * D + d];
            num += e * e;
            den += ref[i * D + d] * ref[i * D + d];
        }
        const double err = den > 0 ? std::sqrt(num / den) : (num > 0 ? INFINITY : 0.0);
        const double bound = band ? std::fmax(1e-5, 6e-7 * kappa[i]) : 1e-12;
        worst = std::fmax(worst, err / bound);
    }
    std::vector<Body<D>> stepped = bodies, want = bodies;
    brute_force_cuda_simulate<D>(stepped, 1.0, 3);
    oracle_simulate(D, n, reinterpret_cast<double*>(want.data()), 4.471e-21, 1e-10, 1.0, 3, 1);
    double xerr = 0.0;
    for (size_t i = 0; i < n; ++i)
        for (int d = 0; d < D; ++d)
            xerr = std::fmax(xerr, std::fabs(stepped[i].position[d] - want[i].position[d]) / 1.0e7);
    std::printf("dim=%d n=%zu worst_force_err_over_bound=%.3f traj_err=%.3e kernel_ms=%.3f\n", D, n, worst, xerr,
                brute_force_cuda_last_kernel_ms());
    return (worst <= 1.0 && xerr <= (band ? 1e-6 : 1e-12)) ? 0 : 1;
}

int main(int argc, char** argv) {
    const int dim = argc > 1 ? std::atoi(argv[1]) : 3;
    const size_t n = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 3000;
    try {
        const int rc = dim == 2 ? run<2>(n) : run<3>(n);
        // error path: exceptions, not crashes (safely_execute catches std::exception, utils.h:95-103)
        bool threw = false;
        setenv("NB200_GPUS", "9999", 1);
        try { std::vector<Body<3>> b(10); brute_force_cuda_n_body<3>(b); } catch (const std::runtime_error&) { threw = true; }
        unsetenv("NB200_GPUS");
        brute_force_cuda_release();
        if (rc == 0 && threw) { std::printf("ADAPTER_OK\n"); return 0; }
        std::printf("ADAPTER_FAIL rc=%d threw=%d\n", rc, (int)threw);
        return 1;
    } catch (const std::exception& e) {
        std::printf("ADAPTER_EXCEPTION %s\n", e.what());
        return 2;
    }
}
