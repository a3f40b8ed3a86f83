// ref_shim.cpp -- extern "C" doorway into the UNMODIFIED reference brute-force code.
//
// TEST INFRASTRUCTURE ONLY (see oracle/nbody_oracle.c header).  This TU is ours; it only
// #includes the reference's headers from where they lie (-I/root/reference/nbody-sim-new)
// and is linked with an object compiled from the reference's own methods.cpp.  No reference
// source is copied into this repository.  Output: oracle/_ref/libnbref.so (git-ignored).
//
// Entry points wrapped (reference file:line):
//   brute_force_seq_n_body<D>       methods.cpp:7-42
//   brute_force_omp_n_body_1<D>     methods.cpp:45-95
//   brute_force_omp_n_body_2<D>     methods.cpp:98-136
//   brute_force_parlay_n_body_1<D>  methods.cpp:139-186
//   brute_force_parlay_n_body_2<D>  methods.cpp:189-224
//   update_body_velocities<D>       methods.cpp:426-438
//   update_body_positions<D>        methods.cpp:441-450
//   compute_accuracy_omp<D>         utils.h:170-219
//
// The reference HEAD does not link as shipped: six FMM member functions are
// declared but never defined.  They are off the brute-force path; the stubs below only exist
// so that the shared object has no undefined symbols (ctypes loads with RTLD_NOW).
#include <chrono>
#include <cstring>
#include <stdexcept>
#include <thread>

#include "methods.h"

static_assert(sizeof(Body<2>) == 40 && sizeof(Body<3>) == 56, "Body<D> AoS stride");
static_assert(sizeof(Vector<2>) == 16 && sizeof(Vector<3>) == 24, "Vector<D> layout");

// ---- stubs for members the reference declares but never defines (fmm.h:98,101; fmm_omp.h:36-45)
template <int D> void FMMNode<D>::translate_local_to_children(int) {
    throw std::logic_error("FMM is outside the brute-force path");
}
template <int D>
void FMMNode<D>::compute_direct_forces(std::vector<Vector<D>>&, const std::vector<Body<D>>&) {
    throw std::logic_error("FMM is outside the brute-force path");
}
template <int D> void FMM_OMP<D>::m2l_phase() { throw std::logic_error("FMM stub"); }
template <int D> void FMM_OMP<D>::l2l_phase() { throw std::logic_error("FMM stub"); }
template <int D>
void FMM_OMP<D>::l2p_phase(std::vector<Vector<D>>&, const std::vector<Body<D>>&) {
    throw std::logic_error("FMM stub");
}
template <int D>
void FMM_OMP<D>::p2p_phase(std::vector<Vector<D>>&, const std::vector<Body<D>>&) {
    throw std::logic_error("FMM stub");
}
#define NB_STUBS(D)                                                                              \
    template void FMMNode<D>::translate_local_to_children(int);                                 \
    template void FMMNode<D>::compute_direct_forces(std::vector<Vector<D>>&,                    \
                                                    const std::vector<Body<D>>&);               \
    template void FMM_OMP<D>::m2l_phase();                                                      \
    template void FMM_OMP<D>::l2l_phase();                                                      \
    template void FMM_OMP<D>::l2p_phase(std::vector<Vector<D>>&, const std::vector<Body<D>>&);  \
    template void FMM_OMP<D>::p2p_phase(std::vector<Vector<D>>&, const std::vector<Body<D>>&);
NB_STUBS(2)
NB_STUBS(3)
#undef NB_STUBS

namespace {

template <int D> std::vector<Body<D>> to_vec(const void* aos, size_t n) {
    std::vector<Body<D>> v(n);
    if (n) std::memcpy(static_cast<void*>(v.data()), aos, n * sizeof(Body<D>));
    return v;
}

template <int D, class Seq> void store(const Seq& f, double* out) {
    for (size_t i = 0; i < f.size(); ++i)
        for (int d = 0; d < D; ++d) out[i * D + d] = f[i][d];
}

// variant: 0 seq, 1 omp_1, 2 omp_2, 3 parlay_1, 4 parlay_2.  Returns seconds spent inside the
// reference call alone (timed the way safely_execute does, utils.h:90-93), or <0 on error.
template <int D> double run(int variant, const void* aos, size_t n, double* out) {
    using clk = std::chrono::high_resolution_clock;
    if (variant <= 2) {
        std::vector<Body<D>> bodies = to_vec<D>(aos, n);
        std::vector<Vector<D>> f;
        auto t0 = clk::now();
        if (variant == 0) f = brute_force_seq_n_body<D>(bodies);
        else if (variant == 1) f = brute_force_omp_n_body_1<D>(bodies);
        else f = brute_force_omp_n_body_2<D>(bodies);
        auto t1 = clk::now();
        if (out) store<D>(f, out);
        return std::chrono::duration<double>(t1 - t0).count();
    }
    std::vector<Body<D>> tmp = to_vec<D>(aos, n);
    parlay::sequence<Body<D>> bodies(tmp.begin(), tmp.end());
    parlay::sequence<Vector<D>> f;
    auto t0 = clk::now();
    if (variant == 3) f = brute_force_parlay_n_body_1<D>(bodies);
    else f = brute_force_parlay_n_body_2<D>(bodies);
    auto t1 = clk::now();
    if (out) store<D>(f, out);
    return std::chrono::duration<double>(t1 - t0).count();
}

// variant as in run(): the force evaluation of every step uses that reference entry point
template <int D> void step(void* aos, size_t n, double dt, int nsteps, int variant) {
    std::vector<Body<D>> bodies = to_vec<D>(aos, n);
    for (int s = 0; s < nsteps; ++s) {
        std::vector<Vector<D>> f;
        if (variant == 0) f = brute_force_seq_n_body<D>(bodies);
        else if (variant == 1) f = brute_force_omp_n_body_1<D>(bodies);
        else if (variant == 2) f = brute_force_omp_n_body_2<D>(bodies);
        else {
            parlay::sequence<Body<D>> pb(bodies.begin(), bodies.end());
            parlay::sequence<Vector<D>> pf =
                variant == 3 ? brute_force_parlay_n_body_1<D>(pb) : brute_force_parlay_n_body_2<D>(pb);
            f.assign(pf.begin(), pf.end());
        }
        update_body_velocities<D>(bodies, f, dt);
        update_body_positions<D>(bodies, dt);
    }
    if (n) std::memcpy(aos, static_cast<const void*>(bodies.data()), n * sizeof(Body<D>));
}

}  // namespace

extern "C" {

// forces_out may be NULL (timing only).  Returns seconds inside the reference call, <0 on error.
double ref_brute_force(int dim, int variant, size_t n, const void* bodies_aos, double* forces_out) {
    try {
        if (variant < 0 || variant > 4) return -1.0;
        if (dim == 2) return run<2>(variant, bodies_aos, n, forces_out);
        if (dim == 3) return run<3>(variant, bodies_aos, n, forces_out);
        return -1.0;
    } catch (...) {
        return -2.0;
    }
}

// nsteps of {force; update_body_velocities; update_body_positions} in place on the AoS buffer.
int ref_simulate(int dim, size_t n, void* bodies_aos, double dt, int nsteps, int variant) {
    try {
        if (variant < 0 || variant > 4) return -1;
        if (dim == 2) step<2>(bodies_aos, n, dt, nsteps, variant);
        else if (dim == 3) step<3>(bodies_aos, n, dt, nsteps, variant);
        else return -1;
        return 0;
    } catch (...) {
        return -2;
    }
}

double ref_accuracy_pct(int dim, size_t n, const double* forces, const double* reference) {
    if (dim == 2) {
        std::vector<Vector<2>> a(n), b(n);
        std::memcpy(static_cast<void*>(a.data()), forces, n * sizeof(Vector<2>));
        std::memcpy(static_cast<void*>(b.data()), reference, n * sizeof(Vector<2>));
        return compute_accuracy_omp<2>(a, b);
    }
    std::vector<Vector<3>> a(n), b(n);
    std::memcpy(static_cast<void*>(a.data()), forces, n * sizeof(Vector<3>));
    std::memcpy(static_cast<void*>(b.data()), reference, n * sizeof(Vector<3>));
    return compute_accuracy_omp<3>(a, b);
}

// The reference's own leaf (P2P) loop: BVH<D>::calculate_force on a tree whose root IS a leaf (max_bodies >= n), i.e.
// the direct sum of bvh.cpp:149-177 over all bodies, in body order, with the reference's guards.  forces_out = n * dim.
int ref_bvh_single_leaf_forces(int dim, size_t n, const void* bodies_aos, double* forces_out) {
    try {
        if (dim == 2) {
            std::vector<Body<2>> b = to_vec<2>(bodies_aos, n);
            BVH<2> tree(b, (int)n + 1);
            store<2>(tree.calculate_forces(b), forces_out);
        } else if (dim == 3) {
            std::vector<Body<3>> b = to_vec<3>(bodies_aos, n);
            BVH<3> tree(b, (int)n + 1);
            store<3>(tree.calculate_forces(b), forces_out);
        } else return -1;
        return 0;
    } catch (...) {
        return -2;
    }
}

// The reference's BVH method end to end (default leaf size 16, opening criterion and all): for timing beside the leaf step.
double ref_bvh_forces(int dim, size_t n, const void* bodies_aos, double* forces_out) {
    using clk = std::chrono::high_resolution_clock;
    try {
        if (dim == 2) {
            std::vector<Body<2>> b = to_vec<2>(bodies_aos, n);
            auto t0 = clk::now();
            std::vector<Vector<2>> f = bvh_seq_n_body<2>(b);
            auto t1 = clk::now();
            if (forces_out) store<2>(f, forces_out);
            return std::chrono::duration<double>(t1 - t0).count();
        }
        std::vector<Body<3>> b = to_vec<3>(bodies_aos, n);
        auto t0 = clk::now();
        std::vector<Vector<3>> f = bvh_seq_n_body<3>(b);
        auto t1 = clk::now();
        if (forces_out) store<3>(f, forces_out);
        return std::chrono::duration<double>(t1 - t0).count();
    } catch (...) {
        return -2.0;
    }
}

double ref_G(void) { return G; }  // utils.h:21
int ref_omp_threads(void) { return omp_get_max_threads(); }
int ref_parlay_workers(void) { return (int)parlay::num_workers(); }

}  // extern "C"
