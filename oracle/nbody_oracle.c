/*
 * nbody_oracle.c -- CPU restatement of the reference's brute-force path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this file's library, and only as the checker
 * (or the timed CPU baseline), never as the thing shipped.  The CUDA product
 * path (libnb200.so) never links, loads or calls anything here.
 *
 * Parity pinning: the reference ships no golden vectors or tests for this
 * path, so this restatement is pinned against the
 * reference ITSELF: oracle/Makefile compiles the unmodified
 * /root/reference/nbody-sim-new/methods.cpp into oracle/_ref/libnbref.so and
 * tests/test_oracle.py requires bit-identical output between this file and
 * that library on seeded inputs (when /root/reference is present), and
 * against the fixtures oracle/gen_golden.py generated from it
 * (tests/golden/, always).
 *
 * Every function follows the reference's operation ORDER literally (plain
 * C, no FMA contraction: build with -ffp-contract=off and no -march, like the
 * reference Makefile:2 which has no -march either).
 *
 * Body layout = the reference's AoS Body<D> (body.h:7-19): D doubles of
 * position, D doubles of velocity, one double of mass; stride 2*D+1 doubles
 * (40 B for D=2, 56 B for D=3).  Forces = Vector<D> (vector.h:9-12): D doubles.
 *
 * G and the pair cut-off are runtime arguments here; the reference hard-codes
 * G = 4.471e-21 (utils.h:21) and 1e-10 (methods.cpp:24,69,119,160,207).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define STRIDE(D) (2 * (D) + 1)
#define POS(b, i, D) ((b) + (size_t)(i) * STRIDE(D))
#define VEL(b, i, D) ((b) + (size_t)(i) * STRIDE(D) + (D))
#define MASS(b, i, D) ((b)[(size_t)(i) * STRIDE(D) + 2 * (D)])

/*
 * One ordered interaction, literally:
 *   diff    = p_j - p_i                          vector.h:32-36
 *   dist_sq = ((0.0 + d0*d0) + d1*d1) [+ d2*d2]  vector.h:81-85
 *   if (dist_sq < cutoff) skip                   methods.cpp:119
 *   dist    = sqrt(dist_sq); dist_cb = dist_sq*dist        methods.cpp:121-122
 *   force_mag = ((G*m_i)*m_j)/dist_cb            methods.cpp:125
 *   mag = sqrt(dist_sq)  (recomputed)            vector.h:94 -> :88-90
 *   force[d] = (diff[d]/mag)*force_mag           vector.h:96, :46-50, :39-43
 * Returns 0 when the pair is skipped, else 1 with force[] filled.
 */
static inline int pair_force(int D, const double *pi, const double *pj, double mi, double mj,
                             double G, double cutoff, double *force)
{
    double diff[3];
    double dist_sq = 0.0;
    for (int d = 0; d < D; d++) diff[d] = pj[d] - pi[d];
    for (int d = 0; d < D; d++) dist_sq += diff[d] * diff[d];
    if (dist_sq < cutoff) return 0;
    double dist = sqrt(dist_sq);
    double dist_cb = dist_sq * dist;
    double force_mag = G * mi * mj / dist_cb;
    double mag = sqrt(dist_sq);
    if (mag < 1e-10) { /* vector.h:95: normalized() returns the zero vector; unreachable when cutoff >= 1e-20 */
        for (int d = 0; d < D; d++) force[d] = 0.0 * force_mag;   /* vector.h:39-43 (0 * inf = NaN for r = 0) */
        return 1;
    }
    for (int d = 0; d < D; d++) force[d] = (diff[d] / mag) * force_mag;
    return 1;
}

/* brute_force_seq_n_body<D>, methods.cpp:7-42: j>i pairs, forces[j]+=f, forces[i]-=f. */
int oracle_forces_seq(int D, size_t n, const double *bodies, double G, double cutoff, double *forces)
{
    if (D != 2 && D != 3) return -1;
    memset(forces, 0, n * (size_t)D * sizeof(double));
    for (size_t i = 0; i < n; i++) {
        for (size_t j = i + 1; j < n; j++) {
            double f[3];
            if (!pair_force(D, POS(bodies, i, D), POS(bodies, j, D), MASS(bodies, i, D),
                            MASS(bodies, j, D), G, cutoff, f))
                continue;
            for (int d = 0; d < D; d++) forces[j * D + d] += f[d];
            for (int d = 0; d < D; d++) forces[i * D + d] -= f[d];
        }
    }
    return 0;
}

/*
 * brute_force_omp_n_body_2<D>, methods.cpp:98-136 (== brute_force_parlay_n_body_2,
 * :189-224): all ordered pairs, row i private, j ascending, forces[i] -= f.
 * Rows are independent, so the OpenMP schedule cannot change any bit.
 */
int oracle_forces_omp2(int D, size_t n, const double *bodies, double G, double cutoff, double *forces)
{
    if (D != 2 && D != 3) return -1;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        double acc[3] = {0.0, 0.0, 0.0};
        for (size_t j = 0; j < n; j++) {
            double f[3];
            if (i == j) continue;
            if (!pair_force(D, POS(bodies, i, D), POS(bodies, j, D), MASS(bodies, i, D),
                            MASS(bodies, j, D), G, cutoff, f))
                continue;
            for (int d = 0; d < D; d++) acc[d] -= f[d];
        }
        for (int d = 0; d < D; d++) forces[i * D + d] = acc[d];
    }
    return 0;
}

/*
 * The omp_2 row sum for a SUBSET of targets (all sources): what a rank of the
 * target-sharded multi-GPU path owns, and the sampled check used at sizes where
 * the full N^2 CPU pass is too slow.  forces_out is nt*D, in targets[] order.
 */
int oracle_forces_targets(int D, size_t n, const double *bodies, double G, double cutoff,
                          const int64_t *targets, size_t nt, double *forces_out)
{
    if (D != 2 && D != 3) return -1;
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t t = 0; t < nt; t++) {
        size_t i = (size_t)targets[t];
        double acc[3] = {0.0, 0.0, 0.0};
        for (size_t j = 0; j < n; j++) {
            double f[3];
            if (i == j) continue;
            if (!pair_force(D, POS(bodies, i, D), POS(bodies, j, D), MASS(bodies, i, D),
                            MASS(bodies, j, D), G, cutoff, f))
                continue;
            for (int d = 0; d < D; d++) acc[d] -= f[d];
        }
        for (int d = 0; d < D; d++) forces_out[t * D + d] = acc[d];
    }
    return 0;
}

/*
 * Same row sums in long double with the closed form F_i = -G m_i sum_j m_j d/r^4
 * ("truth" for sampled targets; SURVEY.md section 4 lesson iii).  The cut-off
 * decision is still taken on the DOUBLE dist_sq so the pair set is identical.
 */
int oracle_forces_targets_ld(int D, size_t n, const double *bodies, double G, double cutoff,
                             const int64_t *targets, size_t nt, double *forces_out)
{
    if (D != 2 && D != 3) return -1;
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t t = 0; t < nt; t++) {
        size_t i = (size_t)targets[t];
        const double *pi = POS(bodies, i, D);
        long double acc[3] = {0.0L, 0.0L, 0.0L};
        for (size_t j = 0; j < n; j++) {
            if (i == j) continue;
            const double *pj = POS(bodies, j, D);
            double dsq = 0.0;
            long double r2 = 0.0L, diff[3];
            for (int d = 0; d < D; d++) {
                double dd = pj[d] - pi[d];
                dsq += dd * dd;
                diff[d] = (long double)pj[d] - (long double)pi[d];
                r2 += diff[d] * diff[d];
            }
            if (dsq < cutoff) continue;
            long double s = (long double)MASS(bodies, j, D) / (r2 * r2);
            for (int d = 0; d < D; d++) acc[d] += s * diff[d];
        }
        long double gm = (long double)G * (long double)MASS(bodies, i, D);
        for (int d = 0; d < D; d++) forces_out[t * D + d] = (double)(-gm * acc[d]);
    }
    return 0;
}

/* update_body_velocities<D>, methods.cpp:426-438:  v[d] += (F[d] / m) * dt. */
int oracle_update_velocities(int D, size_t n, double *bodies, const double *forces, double dt)
{
    if (D != 2 && D != 3) return -1;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        double m = MASS(bodies, i, D);
        double *v = VEL(bodies, i, D);
        for (int d = 0; d < D; d++) v[d] += (forces[i * D + d] / m) * dt;
    }
    return 0;
}

/* update_body_positions<D>, methods.cpp:441-450:  x[d] += v[d] * dt  (the NEW v). */
int oracle_update_positions(int D, size_t n, double *bodies, double dt)
{
    if (D != 2 && D != 3) return -1;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        double *x = POS(bodies, i, D);
        const double *v = VEL(bodies, i, D);
        for (int d = 0; d < D; d++) x[d] += v[d] * dt;
    }
    return 0;
}

/*
 * One "step" as SURVEY.md section 3.3 assembles it from the reference's pieces:
 * F = brute force; update_body_velocities; update_body_positions.
 * variant 0 = seq pair ordering (methods.cpp:7-42), 1 = omp_2 ordering (:98-136).
 */
int oracle_simulate(int D, size_t n, double *bodies, double G, double cutoff, double dt, int nsteps,
                    int variant)
{
    if (D != 2 && D != 3) return -1;
    double *forces = (double *)malloc((n ? n : 1) * (size_t)D * sizeof(double));
    if (!forces) return -2;
    for (int s = 0; s < nsteps; s++) {
        if (variant == 0)
            oracle_forces_seq(D, n, bodies, G, cutoff, forces);
        else
            oracle_forces_omp2(D, n, bodies, G, cutoff, forces);
        oracle_update_velocities(D, n, bodies, forces, dt);
        oracle_update_positions(D, n, bodies, dt);
    }
    free(forces);
    return 0;
}

/*
 * Conserved energy of the reference's own force law:
 * the brute-force path is a REPULSIVE central force of magnitude G m_i m_j / r^3,
 * whose pair potential is G m_i m_j / (2 r^2).  Pairs under the cut-off exert no
 * force and carry no potential.  kinetic = sum 1/2 m v^2.  (New harness
 * behaviour: the reference has no energy diagnostic.)
 */
int oracle_energy(int D, size_t n, const double *bodies, double G, double cutoff, double *kinetic,
                  double *potential)
{
    if (D != 2 && D != 3) return -1;
    double ke = 0.0, pe = 0.0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : ke, pe)
    for (size_t i = 0; i < n; i++) {
        const double *v = VEL(bodies, i, D);
        double v2 = 0.0;
        for (int d = 0; d < D; d++) v2 += v[d] * v[d];
        ke += 0.5 * MASS(bodies, i, D) * v2;
        double row = 0.0;
        for (size_t j = i + 1; j < n; j++) {
            double dsq = 0.0;
            for (int d = 0; d < D; d++) {
                double dd = POS(bodies, j, D)[d] - POS(bodies, i, D)[d];
                dsq += dd * dd;
            }
            if (dsq < cutoff) continue;
            row += MASS(bodies, j, D) / dsq;
        }
        pe += 0.5 * G * MASS(bodies, i, D) * row;
    }
    *kinetic = ke;
    *potential = pe;
    return 0;
}


/*
 * Summation condition number of each body's force (harness diagnostic, not in the reference):
 *   kappa_i = sum_j ||f_ij||_2 / ||sum_j f_ij||_2   >= 1
 * The forward error of ANY finite-precision evaluation of the row sum is bounded by
 * ~ c * u * kappa_i (u = unit roundoff), so FP32-mode parity is stated relative to kappa_i.
 */
int oracle_condition(int D, size_t n, const double *bodies, double G, double cutoff, double *kappa)
{
    if (D != 2 && D != 3) return -1;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        double acc[3] = {0.0, 0.0, 0.0}, sum_abs = 0.0;
        for (size_t j = 0; j < n; j++) {
            double f[3], nf = 0.0;
            if (i == j) continue;
            if (!pair_force(D, POS(bodies, i, D), POS(bodies, j, D), MASS(bodies, i, D),
                            MASS(bodies, j, D), G, cutoff, f))
                continue;
            for (int d = 0; d < D; d++) { acc[d] -= f[d]; nf += f[d] * f[d]; }
            sum_abs += sqrt(nf);
        }
        double na = 0.0;
        for (int d = 0; d < D; d++) na += acc[d] * acc[d];
        na = sqrt(na);
        kappa[i] = na > 0.0 ? sum_abs / na : (sum_abs > 0.0 ? INFINITY : 1.0);
    }
    return 0;
}

/* kappa of a subset of targets (all sources): lets the FP32 criterion be applied to sampled targets
 * at sizes where the full O(N^2) pass above is out of reach. */
int oracle_condition_targets(int D, size_t n, const double *bodies, double G, double cutoff,
                             const long long *targets, size_t ntargets, double *kappa)
{
    if (D != 2 && D != 3) return -1;
    for (size_t k = 0; k < ntargets; k++)
        if (targets[k] < 0 || (size_t)targets[k] >= n) return -2;
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t k = 0; k < ntargets; k++) {
        const size_t i = (size_t)targets[k];
        double acc[3] = {0.0, 0.0, 0.0}, sum_abs = 0.0;
        for (size_t j = 0; j < n; j++) {
            double f[3], nf = 0.0;
            if (i == j) continue;
            if (!pair_force(D, POS(bodies, i, D), POS(bodies, j, D), MASS(bodies, i, D),
                            MASS(bodies, j, D), G, cutoff, f))
                continue;
            for (int d = 0; d < D; d++) { acc[d] -= f[d]; nf += f[d] * f[d]; }
            sum_abs += sqrt(nf);
        }
        double na = 0.0;
        for (int d = 0; d < D; d++) na += acc[d] * acc[d];
        na = sqrt(na);
        kappa[k] = na > 0.0 ? sum_abs / na : (sum_abs > 0.0 ? INFINITY : 1.0);
    }
    return 0;
}

/*
 * The leaf (P2P) branch of the tree codes, restated literally:
 *   BVH<D>::calculate_force, node->is_leaf       bvh.cpp:149-177   eps_same = 1e-9, cutoff = 1e-9, skip_same_index = 0
 *   FMM<D>::calculate_accurate_force, leaf       fmm.cpp:622-637   eps_same < 0,    cutoff = 1e-10, skip_same_index = 1
 * For every body i of target leaf l and every body j of the leaves nbr_leaves[nbr_off[l] .. nbr_off[l+1]), in list order:
 *   same_position = all_d |p_i[d] - p_j[d]| <= eps_same  -> skip      bvh.cpp:156-163
 *   diff = p_j - p_i; dist_sq = sum diff^2 (from 0.0, d ascending)    bvh.cpp:166-167, vector.h:81-85
 *   if (dist_sq < cutoff) skip                                         bvh.cpp:170
 *   dist = sqrt(dist_sq); force_mag = G * m_i * m_j / (dist_sq * dist) bvh.cpp:172-173
 *   force += diff.normalized() * force_mag                             bvh.cpp:175  (attractive; sign = -1 flips it)
 * forces of bodies that sit in no leaf stay zero.
 */
int oracle_p2p_leaves(int D, size_t n, const double *bodies, size_t n_leaves, const long long *leaf_off,
                      const long long *leaf_bodies, const long long *nbr_off, const long long *nbr_leaves, double G,
                      double cutoff, double eps_same, int skip_same_index, int sign, double *forces)
{
    if (D != 2 && D != 3) return -1;
    for (size_t k = 0; k < n * (size_t)D; k++) forces[k] = 0.0;
#pragma omp parallel for schedule(dynamic, 8)
    for (size_t l = 0; l < n_leaves; l++) {
        for (long long a = leaf_off[l]; a < leaf_off[l + 1]; a++) {
            const long long i = leaf_bodies[a];
            const double *pi = POS(bodies, i, D);
            double f[3] = {0.0, 0.0, 0.0};
            for (long long q = nbr_off[l]; q < nbr_off[l + 1]; q++) {
                const long long sl = nbr_leaves[q];
                for (long long b = leaf_off[sl]; b < leaf_off[sl + 1]; b++) {
                    const long long j = leaf_bodies[b];
                    const double *pj = POS(bodies, j, D);
                    if (skip_same_index && j == i) continue;
                    if (eps_same >= 0.0) {
                        int same = 1;
                        for (int d = 0; d < D; d++)
                            if (fabs(pi[d] - pj[d]) > eps_same) { same = 0; break; }
                        if (same) continue;
                    }
                    double diff[3], dist_sq = 0.0;
                    for (int d = 0; d < D; d++) diff[d] = pj[d] - pi[d];
                    for (int d = 0; d < D; d++) dist_sq += diff[d] * diff[d];
                    if (dist_sq < cutoff) continue;
                    const double dist = sqrt(dist_sq);
                    const double force_mag = G * MASS(bodies, i, D) * MASS(bodies, j, D) / (dist_sq * dist);
                    const double mag = sqrt(dist_sq);
                    if (mag < 1e-10) continue;                    /* normalized() -> zero vector: adds 0 (vector.h:95) */
                    for (int d = 0; d < D; d++) f[d] += (diff[d] / mag) * force_mag;
                }
            }
            for (int d = 0; d < D; d++) forces[(size_t)i * D + d] = sign < 0 ? -f[d] : f[d];
        }
    }
    return 0;
}

/* compute_accuracy_omp, utils.h:170-219: % of bodies whose every component is within
 * 1 % of the reference force (|ref| < 1e-20 -> absolute test |f| <= 1e-9). */
double oracle_accuracy_pct(int D, size_t n, const double *forces, const double *ref)
{
    size_t ok = 0;
    for (size_t i = 0; i < n; i++) {
        int good = 1;
        for (int d = 0; d < D; d++) {
            double r = ref[i * D + d], f = forces[i * D + d];
            if (fabs(r) < 1e-20) {
                if (fabs(f) > 1e-9) { good = 0; break; }
                continue;
            }
            if (fabs((f - r) / r) > 0.01) { good = 0; break; }
        }
        ok += (size_t)good;
    }
    return n ? 100.0 * (double)ok / (double)n : 0.0;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun pins OMP_NUM_THREADS=1 for its workers; the checker on rank 0 may take more threads back */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
