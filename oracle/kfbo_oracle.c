/*
 * kfbo_oracle.c -- CPU ORACLE for the kf_bayesopt Monte Carlo cost evaluation.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke()
 * check in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py may load it, and only as the checker / timed CPU baseline.  The
 * product path (kf_bayesopt_b200/ + include/kfbo.h) never links or calls it.
 *
 * WHAT PINS IT.  The reference (arpg/kf_bayesopt, mounted at /root/reference) ships no tests,
 * fixtures or golden vectors for this path (CMakeLists.txt:221-232 is commented out) and seeds
 * every RNG from the wall clock (system_models.hpp:74,147; simulator.cpp:9), so there is nothing
 * of its own to check against -- by that measure parity is UNPINNED.  Its dependencies (Eigen3,
 * ROS, Boost, bayesopt) are absent here, so its build cannot run either.  What this repo does
 * instead: oracle/Makefile compiles the reference's OWN hot-path sources (trial.cpp, simulator.cpp,
 * kalman_filter.cpp, read_para_vehicle.cpp, the headers under include/) from where they lie into
 * oracle/_ref/libkfbo_ref.so, with oracle/ref_shim/ standing in for the few pieces of Eigen3 / ROS /
 * matplotlibcpp they use and a counter for the wall clock.  This file is checked against that
 * library (tests/test_oracle_vs_ref.py: models, Van Loan discretisation, KalmanFilter::Step,
 * Trial::Run end to end with the reference's RNG discipline replayed) and against vectors it
 * produced (tests/golden/ref_vectors.npz, for machines without the reference tree).  What stays
 * restated rather than linked is Eigen's internal arithmetic for 1x1 / 2x2 / 4x4 matrices (product
 * order, closed-form inverse, expm, the eigen-solver's sign convention) -- see ref_shim/Eigen/Dense.
 * Independent pins besides: a numpy/scipy restatement (tests/golden/make_golden.py) and
 * data-independent known answers.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Compile with -O2 -ffp-contract=off: the reference's catkin
 * build targets baseline x86-64 (no FMA), so no contraction here either.
 *
 * Conventions: N=2 states, Me=1 observation (include/system_models.hpp:11-12);
 * 2x2 matrices are row-major double[4] = {m00, m01, m10, m11}.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KO_U_TABLE 2000 /* include/system_models.hpp:51 -- hard limit on steps */

/* ------------------------------------------------------------------------- */
/* 4x4 matrix exponential.  The reference calls Eigen's MatrixBase::exp()     */
/* (include/continuous2discrete.hpp:25; Eigen is a third-party dependency     */
/* that is absent here, unpinned beyond >=3.3 in CMakeLists.txt:21).  Eigen   */
/* uses scaling-and-squaring Pade; we restate the published definition        */
/* exp(A) = sum A^k/k! with scaling-and-squaring.  For the robot1d model the  */
/* Van Loan matrix is nilpotent (A^4 = 0) so the series terminates and both   */
/* agree to rounding (checked against scipy.linalg.expm in tests).            */
/* ------------------------------------------------------------------------- */
static void mat4_mul(const double *a, const double *b, double *c)
{
    double t[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += a[4 * i + k] * b[4 * k + j];
            t[4 * i + j] = s;
        }
    memcpy(c, t, sizeof t);
}

void ko_expm4(const double *A, double *E)
{
    double norm1 = 0.0;
    for (int j = 0; j < 4; ++j) {
        double cs = 0.0;
        for (int i = 0; i < 4; ++i) cs += fabs(A[4 * i + j]);
        if (cs > norm1) norm1 = cs;
    }
    int s = 0;
    double scale = 1.0;
    while (norm1 * scale > 0.5 && s < 60) { scale *= 0.5; ++s; }
    double As[16], term[16], sum[16];
    for (int i = 0; i < 16; ++i) As[i] = A[i] * scale;
    for (int i = 0; i < 16; ++i) sum[i] = term[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int k = 1; k <= 40; ++k) {
        mat4_mul(term, As, term);
        double mx = 0.0;
        for (int i = 0; i < 16; ++i) { term[i] /= (double)k; if (fabs(term[i]) > mx) mx = fabs(term[i]); }
        for (int i = 0; i < 16; ++i) sum[i] += term[i];
        if (mx == 0.0) break; /* nilpotent: series terminated exactly */
    }
    for (int i = 0; i < s; ++i) mat4_mul(sum, sum, sum);
    memcpy(E, sum, sizeof sum);
}

/* include/continuous2discrete.hpp:8-32 with the Fc, Qc of
 * include/system_models.hpp:33-39.  Qc(0,0), Qc(0,1), Qc(1,0) are left
 * uninitialised by the reference (only (1,1) is assigned); the intended
 * white-noise-acceleration model has them 0 and that is what is restated. */
void ko_discretize(double q, double dt, double *Fd, double *Qd)
{
    const double Fc[4] = {0.0, 1.0, 0.0, 0.0};
    const double Qc[4] = {0.0, 0.0, 0.0, q};
    double bigA[16], bigB[16];
    memset(bigA, 0, sizeof bigA);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            bigA[4 * i + j] = -Fc[2 * i + j] * dt;            /* topLeft  = -Fc*dt   (:18) */
            bigA[4 * i + (j + 2)] = Qc[2 * i + j] * dt;        /* topRight = Qc*dt    (:19) */
            bigA[4 * (i + 2) + (j + 2)] = Fc[2 * j + i] * dt;  /* botRight = Fc^T*dt  (:21) */
        }
    ko_expm4(bigA, bigB);
    /* Fd = bottomRight^T (:29) */
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) Fd[2 * i + j] = bigB[4 * (j + 2) + (i + 2)];
    /* Qd = Fd * topRight (:30) */
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            Qd[2 * i + j] = Fd[2 * i + 0] * bigB[4 * 0 + (j + 2)] + Fd[2 * i + 1] * bigB[4 * 1 + (j + 2)];
}

/* include/system_models.hpp:49-56 -- u_[i] = 2cos(0.75 j), j accumulated by j += dt. */
void ko_control_table(double dt, int n, double *u)
{
    double j = 0.0;
    for (int i = 0; i < n; ++i) {
        u[i] = 2 * cos(0.75 * j);
        j += dt;
    }
}

/* One model pair as the reference builds it: ProcessModel (system_models.hpp:30-78)
 * and ObservationModel (:137-150). */
typedef struct {
    double Fd[4], Qd[4], g[2], H[2], R;
    double u[KO_U_TABLE];
} ko_model;

void ko_model_init(ko_model *m, double pnoise0, double onoise0, double dt)
{
    m->g[0] = 0.5 * dt * dt; /* :46 */
    m->g[1] = dt;            /* :47 */
    ko_control_table(dt, KO_U_TABLE, m->u);
    ko_discretize(pnoise0, dt, m->Fd, m->Qd); /* :58 */
    m->H[0] = 1;             /* :141 */
    m->H[1] = 0;             /* :142 */
    m->R = onoise0 / dt;     /* :143 */
}

/* ProcessModel::Predict without the sampled noise (system_models.hpp:82):
 * xpred = fd_*x + g_*u_[count]. */
static inline void process_predict(const ko_model *m, const double *x, int count, double *xp)
{
    const double u = m->u[count];
    xp[0] = (m->Fd[0] * x[0] + m->Fd[1] * x[1]) + m->g[0] * u;
    xp[1] = (m->Fd[2] * x[0] + m->Fd[3] * x[1]) + m->g[1] * u;
}

typedef struct {
    double xest[2], pest[4], nu, sob;
} ko_kf;

/* KalmanFilter::Step, kalman_filter.cpp:18-38 (operation order preserved:
 * (F*P)*F^T + Q; K = c * (1/S); P = (I - K*H) * Ppred, no symmetrisation). */
static inline void kf_step(const ko_model *m, ko_kf *f, double z, int count)
{
    double xpred[2], FP[4], ppred[4], c[2], k[2], X[4];
    process_predict(m, f->xest, count, xpred); /* :21-22 */
    const double *F = m->Fd, *P = f->pest;
    FP[0] = F[0] * P[0] + F[1] * P[2];
    FP[1] = F[0] * P[1] + F[1] * P[3];
    FP[2] = F[2] * P[0] + F[3] * P[2];
    FP[3] = F[2] * P[1] + F[3] * P[3];
    ppred[0] = (FP[0] * F[0] + FP[1] * F[1]) + m->Qd[0]; /* :23 */
    ppred[1] = (FP[0] * F[2] + FP[1] * F[3]) + m->Qd[1];
    ppred[2] = (FP[2] * F[0] + FP[3] * F[1]) + m->Qd[2];
    ppred[3] = (FP[2] * F[2] + FP[3] * F[3]) + m->Qd[3];
    c[0] = ppred[0] * m->H[0] + ppred[1] * m->H[1]; /* :26 */
    c[1] = ppred[2] * m->H[0] + ppred[3] * m->H[1];
    f->sob = (m->H[0] * c[0] + m->H[1] * c[1]) + m->R; /* :27 */
    const double sinv = 1.0 / f->sob;                   /* :28 sob_.inverse() of a 1x1 */
    k[0] = c[0] * sinv;
    k[1] = c[1] * sinv;
    f->nu = z - (m->H[0] * xpred[0] + m->H[1] * xpred[1]); /* :31 */
    f->xest[0] = xpred[0] + k[0] * f->nu;                  /* :33 */
    f->xest[1] = xpred[1] + k[1] * f->nu;
    X[0] = 1.0 - k[0] * m->H[0]; /* :36 */
    X[1] = 0.0 - k[0] * m->H[1];
    X[2] = 0.0 - k[1] * m->H[0];
    X[3] = 1.0 - k[1] * m->H[1];
    f->pest[0] = X[0] * ppred[0] + X[1] * ppred[2]; /* :37 */
    f->pest[1] = X[0] * ppred[1] + X[1] * ppred[3];
    f->pest[2] = X[2] * ppred[0] + X[3] * ppred[2];
    f->pest[3] = X[2] * ppred[1] + X[3] * ppred[3];
}

/* trial.cpp:53-60.  Eigen's fixed 2x2 inverse is adjugate * (1/det);
 * (e^T * Pinv) is formed first, then * e. */
static inline void chi2(const double *x, const ko_kf *f, double *nees, double *nis)
{
    const double e0 = x[0] - f->xest[0], e1 = x[1] - f->xest[1];
    const double *P = f->pest;
    const double det = P[0] * P[3] - P[2] * P[1];
    const double invdet = 1.0 / det;
    const double i00 = P[3] * invdet, i01 = -P[1] * invdet, i10 = -P[2] * invdet, i11 = P[0] * invdet;
    const double r0 = e0 * i00 + e1 * i10, r1 = e0 * i01 + e1 * i11;
    *nees = r0 * e0 + r1 * e1;
    const double sinv = 1.0 / f->sob;
    *nis = (f->nu * sinv) * f->nu;
}

/* Body of one trial, trial.cpp:31-98, with the noise handed in post-transform:
 * x0n = draw of N(0,P0) (simulator.cpp:12), w(step) = draw of N(0,Qd_truth)
 * (system_models.hpp:87,96), v(step) = draw of N(0,R_truth) (:158).
 * out = {row_nees, row_nis, var_nees, var_nis}.  his_* is caller scratch [T]. */
typedef void (*ko_noise_fn)(void *ctx, int step, double *w0, double *w1, double *v);

static void run_trial(const ko_model *truth, const ko_model *est, int T, const double *x0n,
                      ko_noise_fn noise, void *ctx, double *his_nees, double *his_nis, double *out)
{
    double x[2] = {0.0 + x0n[0], 0.0 + x0n[1]}; /* trial.cpp:31, simulator.cpp:12 */
    ko_kf f;
    f.xest[0] = 0.0; f.xest[1] = 0.0;           /* kalman_filter.cpp:14 */
    f.pest[0] = 1.0; f.pest[1] = 0.0; f.pest[2] = 0.0; f.pest[3] = 1.0; /* trial.cpp:32 */
    f.nu = 0.0; f.sob = 1.0;
    double row_nees = 0.0, row_nis = 0.0;
    for (int step = 0; step < T; ++step) {
        double w0, w1, v, xp[2];
        noise(ctx, step, &w0, &w1, &v);
        process_predict(truth, x, step, xp); /* simulator.cpp:17 */
        x[0] = xp[0] + w0;                   /* system_models.hpp:96 */
        x[1] = xp[1] + w1;
        const double z = (truth->H[0] * x[0] + truth->H[1] * x[1]) + v; /* simulator.cpp:18 */
        kf_step(est, &f, z, step);           /* trial.cpp:50 */
        double nees, nis;
        chi2(x, &f, &nees, &nis);
        row_nees += nees;                    /* trial.cpp:62-66 */
        row_nis += nis;
        his_nees[step] = nees;
        his_nis[step] = nis;
    }
    const double ave_nees = row_nees / T, ave_nis = row_nis / T; /* :84-85 */
    double var_nees = 0.0, var_nis = 0.0;
    for (int i = 0; i < T; ++i) {                                  /* :90-93 */
        var_nees += (his_nees[i] - ave_nees) * (his_nees[i] - ave_nees);
        var_nis += (his_nis[i] - ave_nis) * (his_nis[i] - ave_nis);
    }
    var_nees /= T - 1; /* :95-96 */
    var_nis /= T - 1;
    out[0] = row_nees; out[1] = row_nis; out[2] = var_nees; out[3] = var_nis;
}

/* ------------------------------------------------------------------------- */
/* Supplied-noise evaluation: the exact-parity entry point.                    */
/* Layouts (SoA, trial index fastest): x0noise[2][B], w[T][2][B], v[T][B],     */
/* per_trial[4][B] = {row_nees, row_nis, var_nees, var_nis}.                   */
/* ------------------------------------------------------------------------- */
typedef struct { const double *w, *v; long B, i; } supplied_ctx;

static void supplied_noise(void *p, int step, double *w0, double *w1, double *v)
{
    const supplied_ctx *c = (const supplied_ctx *)p;
    *w0 = c->w[((long)step * 2 + 0) * c->B + c->i];
    *w1 = c->w[((long)step * 2 + 1) * c->B + c->i];
    *v = c->v[(long)step * c->B + c->i];
}

int ko_eval_supplied(double dt, int T, double q_est, double r_est, long B, const double *x0noise,
                     const double *w, const double *v, double *per_trial)
{
    if (T < 2 || T > KO_U_TABLE || B < 0) return -1;
    ko_model *truth = malloc(sizeof *truth), *est = malloc(sizeof *est);
    double *hn = malloc(sizeof(double) * T), *hi = malloc(sizeof(double) * T);
    /* truth noise is supplied, so its (q, r) never enter; Fd, g, u, H do */
    ko_model_init(truth, 1.0, 1.0, dt);
    ko_model_init(est, q_est, r_est, dt);
    for (long i = 0; i < B; ++i) {
        supplied_ctx c = {w, v, B, i};
        const double x0n[2] = {x0noise[i], x0noise[B + i]};
        double out[4];
        run_trial(truth, est, T, x0n, supplied_noise, &c, hn, hi, out);
        for (int k = 0; k < 4; ++k) per_trial[(long)k * B + i] = out[k];
    }
    free(truth); free(est); free(hn); free(hi);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as     */
/* easy as 1, 2, 3", SC'11; Random123 v1.14).  Not part of the reference --    */
/* it replaces eigenmvn.h's racy static mt19937 as north_star prescribes.      */
/* The stream layout below is the product's on-device layout;     */
/* restated here so the device RNG path can be checked draw for draw.          */
/* ------------------------------------------------------------------------- */
void ko_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out)
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 64 random bits (w0, w1) -> two independent standard normals (Box-Muller), the product's spec
 * (kf_bayesopt_b200/csrc/kfbo_math.cuh: normal_pair), evaluated here with libm:
 *   radius   : 44 bits  m = (w0 << 12) | (w1 >> 20);  u1 = 1 - (m + 1/2) 2^-44;  rad = sqrt(-2 ln u1)
 *   direction: 20 bits  dir = w1 & 0xFFFFF;           theta = 2 pi (dir + 1/2) / 2^20
 *   n_a = rad cos(theta), n_b = rad sin(theta)
 * theta is reduced to its octant exactly (dir + 1/2 and 2^17 - it are exact) so libm sees |arg| <= pi/4. */
void ko_normal_pair(uint32_t w0, uint32_t w1, double *na, double *nb)
{
    const uint64_t m = ((uint64_t)w0 << 12) | (w1 >> 20);
    const double u1 = 1.0 - ((double)m + 0.5) * 0x1p-44;
    const double r = sqrt(-2.0 * log(u1));
    const uint32_t dir = w1 & 0xFFFFFu;
    const unsigned o = dir >> 17;                                  /* octant: 2^17 directions each */
    const double f = ((double)(dir & 0x1FFFFu) + 0.5) * 0x1p-17;   /* position inside the octant, in (0, 1) */
    const double x = (o & 1u) ? 1.0 - f : f;                       /* mirrored for odd octants */
    const double sx = sin(M_PI_4 * x), cx = cos(M_PI_4 * x);
    const double sa = (o & 1u) ? cx : sx, ca = (o & 1u) ? sx : cx; /* angle within the quadrant */
    double sn, cs;
    switch (o >> 1) {
    case 0: sn = sa; cs = ca; break;
    case 1: sn = ca; cs = -sa; break;
    case 2: sn = -sa; cs = -ca; break;
    default: sn = -ca; cs = sa; break;
    }
    *na = r * cs;
    *nb = r * sn;
}

/* The random words of one STEP GROUP (steps 4g .. 4g+3): three Philox calls A, B, C -> six normal pairs:
 *   process noise of step 4g+s: (A.x,A.y), (A.z,A.w), (B.x,B.y), (B.z,B.w) for s = 0..3  -> pn[s][0..1]
 *   observation noise: the two halves of (C.x,C.y) for steps 4g, 4g+1, of (C.z,C.w) for 4g+2, 4g+3 -> on[s] */
void ko_normals_of_group(const uint32_t *A, const uint32_t *B, const uint32_t *Cc, double *pn /*[4][2]*/, double *on /*[4]*/)
{
    ko_normal_pair(A[0], A[1], &pn[0], &pn[1]);
    ko_normal_pair(A[2], A[3], &pn[2], &pn[3]);
    ko_normal_pair(B[0], B[1], &pn[4], &pn[5]);
    ko_normal_pair(B[2], B[3], &pn[6], &pn[7]);
    ko_normal_pair(Cc[0], Cc[1], &on[0], &on[1]);
    ko_normal_pair(Cc[2], Cc[3], &on[2], &on[3]);
}

/* Closed-form Cholesky of the truth Qd and sqrt(R): any T with T T^T = Qd is
 * statistically equivalent to eigenmvn.h:126-130's V*sqrt(Lambda). */
void ko_noise_transform(const double *Qd, double R, double *L /*[3]: L00,L10,L11*/, double *sr)
{
    L[0] = sqrt(Qd[0]);
    L[1] = Qd[2] / L[0];
    L[2] = sqrt(Qd[3] - L[1] * L[1]);
    *sr = sqrt(R);
}

typedef struct {
    uint32_t key[2], trial, stream, stream_x;   /* stream id bits 0..31; bits 32..61 shifted left by 2 */
    double L[3], sr;
    int cached_group;  /* index of the step group whose 12 normals are in pn[], on[] */
    double pn[8], on[4];
} philox_ctx;

static void philox_noise(void *p, int step, double *w0, double *w1, double *v)
{
    philox_ctx *c = (philox_ctx *)p;
    const int group = step >> 2, s = step & 3;
    if (group != c->cached_group) {
        uint32_t out[3][4];
        for (uint32_t j = 0; j < 3; ++j) {
            const uint32_t ctr[4] = {c->trial, (uint32_t)group + 1u, c->stream, j ^ c->stream_x};   /* (trial, block, stream lo, sub ^ stream hi << 2) */
            ko_philox4x32_10(ctr, c->key, out[j]);
        }
        ko_normals_of_group(out[0], out[1], out[2], c->pn, c->on);
        c->cached_group = group;
    }
    const double n0 = c->pn[2 * s], n1 = c->pn[2 * s + 1], n2 = c->on[s];
    *w0 = c->L[0] * n0;
    *w1 = c->L[1] * n0 + c->L[2] * n1;
    *v = c->sr * n2;
}

/* Philox-stream evaluation of trials [trial0, trial0+ntrials) of one
 * (candidate, sample time).  per_trial[4][ntrials]. */
int ko_eval_philox(uint64_t seed, uint64_t stream /* 62-bit id */, double dt, int T, double q_truth, double r_truth,
                   double q_est, double r_est, uint32_t trial0, long ntrials, double *per_trial)
{
    if (T < 2 || T > KO_U_TABLE || ntrials < 0) return -1;
    ko_model *truth = malloc(sizeof *truth), *est = malloc(sizeof *est);
    double *hn = malloc(sizeof(double) * T), *hi = malloc(sizeof(double) * T);
    ko_model_init(truth, q_truth, r_truth, dt);
    ko_model_init(est, q_est, r_est, dt);
    philox_ctx c;
    c.key[0] = (uint32_t)seed; c.key[1] = (uint32_t)(seed >> 32);
    c.stream = (uint32_t)stream;
    c.stream_x = (uint32_t)((stream >> 32) & 0x3FFFFFFFu) << 2;
    ko_noise_transform(truth->Qd, truth->R, c.L, &c.sr);
    for (long i = 0; i < ntrials; ++i) {
        c.trial = trial0 + (uint32_t)i;
        c.cached_group = -1;
        const uint32_t ctr[4] = {c.trial, 0u, c.stream, c.stream_x};              /* block 0, sub 0: the x0 draw */
        uint32_t o[4];
        double x0n[2], out[4];
        ko_philox4x32_10(ctr, c.key, o);
        ko_normal_pair(o[0], o[1], &x0n[0], &x0n[1]); /* P0 = I: transform is the identity */
        run_trial(truth, est, T, x0n, philox_noise, &c, hn, hi, out);
        for (int k = 0; k < 4; ++k) per_trial[(long)k * ntrials + i] = out[k];
    }
    free(truth); free(est); free(hn); free(hi);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Thread averages and cost assembly.                                          */
/* ------------------------------------------------------------------------- */

/* Trial::Run's per-thread reduction, trial.cpp:19,28,81-82,97-98,110-113, for
 * thread `th` of `cores`, given the per-trial outputs of the B' = B/cores trials
 * it ran (per_trial[4][ld], this thread's trials at columns [off, off+B')).
 * avg = {average_nees_, average_nis_, average_var_nees_, average_var_nis_}. */
void ko_thread_averages(const double *per_trial, long ld, long off, int T, int nsimruns, int cores,
                        double *avg)
{
    const double k_average = (double)(T * nsimruns / cores); /* :19 (int arithmetic, then double) */
    const int n = nsimruns / cores;                           /* :28 */
    double all_nees = 0, all_nis = 0, all_var_nees = 0, all_var_nis = 0;
    for (int i = 0; i < n; ++i) {
        all_nees += per_trial[0 * ld + off + i];
        all_nis += per_trial[1 * ld + off + i];
        all_var_nees += per_trial[2 * ld + off + i];
        all_var_nis += per_trial[3 * ld + off + i];
    }
    avg[2] = all_var_nees * cores / nsimruns; /* :110 */
    avg[3] = all_var_nis * cores / nsimruns;  /* :111 */
    avg[0] = all_nees / k_average;            /* :112 */
    avg[1] = all_nis / k_average;             /* :113 */
}

/* trial_node.cpp:70-102: mean of the thread getters, then the cost.
 * cost_choice: 0 JNEES, 1 CNEES, 2 JNIS, 3 CNIS.
 * out = {average_nees, average_nis, average_var_nees, average_var_nis, J}. */
void ko_cost_from_threads(const double *thread_avg /*[cores][4]*/, int cores, int state_dof,
                          int obs_dof, int cost_choice, double *out)
{
    double a[4] = {0, 0, 0, 0};
    for (int i = 0; i < cores; ++i)
        for (int k = 0; k < 4; ++k) a[k] += thread_avg[4 * i + k] / cores; /* :74-75, :90-91 */
    for (int k = 0; k < 4; ++k) out[k] = a[k];
    if (cost_choice == 0 || cost_choice == 1) {
        const double j = fabs(log(a[0] / state_dof));          /* :77 */
        const double jv = fabs(log(0.5 * a[2] / state_dof));   /* :78 */
        out[4] = (cost_choice == 1) ? j + jv : j;               /* :79-82 */
    } else {
        const double j = fabs(log(a[1] / obs_dof));            /* :93 */
        const double jv = fabs(log(0.5 * a[3] / obs_dof));     /* :94 */
        out[4] = (cost_choice == 3) ? j + jv : j;               /* :95-97 */
    }
}

/* trial_node.cpp:118-119: max_element and the FIRST maximal index. */
void ko_max_cost(const double *pic_max, int n, double *max_cost, int *index_max)
{
    int im = 0;
    for (int i = 1; i < n; ++i)
        if (pic_max[im] < pic_max[i]) im = i;
    *max_cost = pic_max[im];
    *index_max = im;
}

/* ------------------------------------------------------------------------- */
/* The threaded CPU path as the reference runs it (trial_node.cpp:54-67): the  */
/* timed baseline.  RNG restates libstdc++'s std::mt19937 +                    */
/* std::normal_distribution<double> (Marsaglia polar with one cached value,    */
/* generate_canonical<double,53> = two 32-bit draws) as eigenmvn.h:51-57 uses  */
/* them, and eigenmvn.h:126-130's transform V*sqrt(max(Lambda,0)).  One        */
/* generator per thread instead of the reference's racy process-wide static   */
/* (eigenmvn.h:51,62); it is re-seeded at every trial like simulator.cpp:9-10. */
/* Heap-free, so faster than the reference's Eigen-dynamic code would be.      */
/* ------------------------------------------------------------------------- */
typedef struct { uint32_t mt[624]; int idx; } mt19937_t;

static void mt_seed(mt19937_t *g, uint32_t s)
{
    g->mt[0] = s;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

static uint32_t mt_next(mt19937_t *g)
{
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            const uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}

uint32_t ko_mt19937_nth(uint32_t seed, int n) /* KAT hook: n-th output (1-based) */
{
    mt19937_t g; mt_seed(&g, seed);
    uint32_t y = 0;
    for (int i = 0; i < n; ++i) y = mt_next(&g);
    return y;
}

static double mt_canonical(mt19937_t *g)
{
    const double lo = (double)mt_next(g), hi = (double)mt_next(g);
    double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}

typedef struct { int has; double saved; } normal_t;

static double mt_normal(mt19937_t *g, normal_t *d)
{
    if (d->has) { d->has = 0; return d->saved; }
    double x, y, r2;
    do {
        x = 2.0 * mt_canonical(g) - 1.0;
        y = 2.0 * mt_canonical(g) - 1.0;
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0.0);
    const double mult = sqrt(-2 * log(r2) / r2);
    d->saved = x * mult; d->has = 1;
    return y * mult;
}

/* 2x2 symmetric eigen-decomposition -> transform = V * diag(sqrt(max(lambda,0))),
 * eigenmvn.h:126-130 (eigenvalues ascending, as SelfAdjointEigenSolver orders). */
void ko_eig_transform2(const double *C, double *Tm)
{
    const double a = C[0], b = 0.5 * (C[1] + C[2]), d = C[3];
    const double tr = a + d, df = a - d;
    const double rt = sqrt(df * df + 4.0 * b * b);
    double l1 = 0.5 * (tr - rt), l2 = 0.5 * (tr + rt); /* ascending */
    double v1[2], v2[2];
    if (b == 0.0) {
        if (a <= d) { v1[0] = 1; v1[1] = 0; v2[0] = 0; v2[1] = 1; l1 = a; l2 = d; }
        else        { v1[0] = 0; v1[1] = 1; v2[0] = 1; v2[1] = 0; l1 = d; l2 = a; }
    } else {
        v2[0] = l2 - d; v2[1] = b;
        double n = sqrt(v2[0] * v2[0] + v2[1] * v2[1]);
        v2[0] /= n; v2[1] /= n;
        v1[0] = -v2[1]; v1[1] = v2[0];
    }
    const double s1 = sqrt(l1 > 0 ? l1 : 0), s2 = sqrt(l2 > 0 ? l2 : 0);
    Tm[0] = v1[0] * s1; Tm[1] = v2[0] * s2;
    Tm[2] = v1[1] * s1; Tm[3] = v2[1] * s2;
}

typedef struct {
    mt19937_t *g; normal_t pn, on;
    double Tq[4], sr;
} mt_ctx;

static void mt_noise(void *p, int step, double *w0, double *w1, double *v)
{
    (void)step;
    mt_ctx *c = (mt_ctx *)p;
    const double n0 = mt_normal(c->g, &c->pn), n1 = mt_normal(c->g, &c->pn); /* samples(1): 2x1 draw */
    *w0 = c->Tq[0] * n0 + c->Tq[1] * n1;
    *w1 = c->Tq[2] * n0 + c->Tq[3] * n1;
    *v = c->sr * mt_normal(c->g, &c->on);
}

typedef struct {
    double dt, q_truth, r_truth, q_est, r_est;
    int T, nsimruns, cores, th;
    uint32_t seed;
    double avg[4];
} mt_job;

static void *mt_thread(void *p)
{
    mt_job *j = (mt_job *)p;
    /* Trial(rv_gt, rv): Simulator(rv_gt), KalmanFilter(rv)  (trial.cpp:12-13) */
    ko_model *truth = malloc(sizeof *truth), *est = malloc(sizeof *est);
    ko_model_init(truth, j->q_truth, j->r_truth, j->dt);
    ko_model_init(est, j->q_est, j->r_est, j->dt);
    mt19937_t *g = malloc(sizeof *g);
    mt_ctx c; c.g = g; c.pn.has = c.on.has = 0;
    ko_eig_transform2(truth->Qd, c.Tq);
    c.sr = sqrt(truth->R > 0 ? truth->R : 0);
    const int n = j->nsimruns / j->cores;
    double *hn = malloc(sizeof(double) * j->T), *hi = malloc(sizeof(double) * j->T);
    double *pt = malloc(sizeof(double) * 4 * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
        mt_seed(g, j->seed + 7919u * (uint32_t)(j->th * n + i)); /* simulator.cpp:9-10 re-seeds per trial */
        normal_t ic = {0, 0.0};
        double x0n[2], out[4];
        x0n[0] = mt_normal(g, &ic); /* P0 = I: V*sqrt(Lambda) = I */
        x0n[1] = mt_normal(g, &ic);
        run_trial(truth, est, j->T, x0n, mt_noise, &c, hn, hi, out);
        for (int k = 0; k < 4; ++k) pt[(long)k * n + i] = out[k];
    }
    ko_thread_averages(pt, n, 0, j->T, j->nsimruns, j->cores, j->avg);
    free(truth); free(est); free(g); free(hn); free(hi); free(pt);
    return NULL;
}

/* One sample time of RunTrial::ProcessNoiseCallback (trial_node.cpp:54-102):
 * `cores` threads x nsimruns/cores trials.  out[5] as ko_cost_from_threads. */
int ko_eval_mt(double dt, int T, double q_truth, double r_truth, double q_est, double r_est,
               int nsimruns, int cores, uint32_t seed, int state_dof, int obs_dof, int cost_choice,
               double *out)
{
    if (T < 2 || T > KO_U_TABLE || cores < 1 || nsimruns < cores) return -1;
    mt_job *jobs = calloc((size_t)cores, sizeof *jobs);
    pthread_t *th = calloc((size_t)cores, sizeof *th);
    for (int i = 0; i < cores; ++i) {
        mt_job j = {dt, q_truth, r_truth, q_est, r_est, T, nsimruns, cores, i, seed, {0, 0, 0, 0}};
        jobs[i] = j;
    }
    for (int i = 0; i < cores; ++i) pthread_create(&th[i], NULL, mt_thread, &jobs[i]); /* :62-63 */
    for (int i = 0; i < cores; ++i) pthread_join(th[i], NULL);                          /* :66-67 */
    double *ta = malloc(sizeof(double) * 4 * (size_t)cores);
    for (int i = 0; i < cores; ++i) memcpy(ta + 4 * i, jobs[i].avg, sizeof(double) * 4);
    ko_cost_from_threads(ta, cores, state_dof, obs_dof, cost_choice, out);
    free(ta); free(jobs); free(th);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Replay of ONE Trial object exactly as the reference's own code runs it when  */
/* its clock is the counter of oracle/ref_shim/fake_clock.h -- the check of     */
/* this file against oracle/_ref (the reference's sources compiled here).       */
/* What the reference does with its static generator (eigenmvn.h:51-62,91-98):  */
/*   * every sampler construction re-seeds it: Trial's constructor reads the    */
/*     clock twice (truth ProcessModel, truth ObservationModel,                 */
/*     system_models.hpp:74,147) and the estimator's default-constructed        */
/*     samplers re-seed it to 5489 -- all overwritten before any draw, because  */
/*   * Simulator::Initialize re-seeds from the clock at the start of EVERY      */
/*     trial (simulator.cpp:9-10): trial i uses seed (unsigned)(clock0 + 2 + i);*/
/*   * samples(n) hands the generator functor to NullaryExpr BY VALUE           */
/*     (eigenmvn.h:135-138): each call starts from a normal_distribution        */
/*     without a saved value, so a 2-vector consumes one polar pair and the     */
/*     1x1 observation draw consumes one pair and DROPS its second half.        */
/* ------------------------------------------------------------------------- */
typedef struct { mt19937_t *g; double Tq[4], sr; } refclock_ctx;

static void refclock_noise(void *p, int step, double *w0, double *w1, double *v)
{
    (void)step;
    refclock_ctx *c = (refclock_ctx *)p;
    normal_t pn = {0, 0.0}, on = {0, 0.0};
    const double n0 = mt_normal(c->g, &pn), n1 = mt_normal(c->g, &pn);   /* system_models.hpp:87 */
    *w0 = c->Tq[0] * n0 + c->Tq[1] * n1;
    *w1 = c->Tq[2] * n0 + c->Tq[3] * n1;
    *v = c->sr * mt_normal(c->g, &on);                                    /* system_models.hpp:158 */
}

int ko_trial_run_refclock(double dt, int T, double q_truth, double r_truth, double q_est, double r_est,
                          int nsimruns, int cores, long long clock0, double *avg4)
{
    if (T < 2 || T > KO_U_TABLE || cores < 1 || nsimruns < cores) return -1;
    ko_model *truth = malloc(sizeof *truth), *est = malloc(sizeof *est);
    ko_model_init(truth, q_truth, r_truth, dt);
    ko_model_init(est, q_est, r_est, dt);
    mt19937_t *g = malloc(sizeof *g);
    refclock_ctx c; c.g = g;
    ko_eig_transform2(truth->Qd, c.Tq);
    c.sr = sqrt(truth->R > 0 ? truth->R : 0);
    const int n = nsimruns / cores;
    double *hn = malloc(sizeof(double) * T), *hi = malloc(sizeof(double) * T);
    double *pt = malloc(sizeof(double) * 4 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        mt_seed(g, (uint32_t)(unsigned long long)(clock0 + 2 + i));      /* simulator.cpp:9-10 */
        normal_t ic = {0, 0.0};
        double x0n[2], out[4];
        x0n[0] = mt_normal(g, &ic);   /* P0 = I: V*sqrt(Lambda) = I (simulator.cpp:10-12) */
        x0n[1] = mt_normal(g, &ic);
        run_trial(truth, est, T, x0n, refclock_noise, &c, hn, hi, out);
        for (int k = 0; k < 4; ++k) pt[(long)k * n + i] = out[k];
    }
    ko_thread_averages(pt, n, 0, T, nsimruns, cores, avg4);
    free(truth); free(est); free(g); free(hn); free(hi); free(pt);
    return 0;
}

/* The post-transform noise draws of trial i of that replay: x0n[2], w[T][2], v[T] -- what
 * Simulator::Initialize / ProcessModel::Predict / ObservationModel::Predict add (simulator.cpp:12,
 * system_models.hpp:96,158).  Feeding them to a supplied-noise evaluation reproduces the trial. */
int ko_refclock_draws(double dt, int T, double q_truth, double r_truth, long long clock0, int i,
                      double *x0n, double *w, double *v)
{
    if (T < 1 || T > KO_U_TABLE) return -1;
    ko_model *truth = malloc(sizeof *truth);
    ko_model_init(truth, q_truth, r_truth, dt);
    mt19937_t *g = malloc(sizeof *g);
    refclock_ctx c; c.g = g;
    ko_eig_transform2(truth->Qd, c.Tq);
    c.sr = sqrt(truth->R > 0 ? truth->R : 0);
    mt_seed(g, (uint32_t)(unsigned long long)(clock0 + 2 + i));
    normal_t ic = {0, 0.0};
    x0n[0] = mt_normal(g, &ic);
    x0n[1] = mt_normal(g, &ic);
    for (int s = 0; s < T; ++s) refclock_noise(&c, s, &w[2 * s], &w[2 * s + 1], &v[s]);
    free(truth); free(g);
    return 0;
}

/* KalmanFilter::Initialize(0, I) then Step(z[k], k) (kalman_filter.cpp:12-38) on a supplied measurement
 * sequence.  out[k] = {xest0, xest1, P00, P01, P10, P11, nu, S}. */
int ko_kf_run(double q_est, double r_est, double dt, int T, const double *z, double *out)
{
    if (T < 1 || T > KO_U_TABLE) return -1;
    ko_model *est = malloc(sizeof *est);
    ko_model_init(est, q_est, r_est, dt);
    ko_kf f;
    f.xest[0] = f.xest[1] = 0; f.pest[0] = 1; f.pest[1] = 0; f.pest[2] = 0; f.pest[3] = 1; f.nu = 0; f.sob = 1;
    for (int k = 0; k < T; ++k) {
        kf_step(est, &f, z[k], k);
        double *o = out + 8 * k;
        o[0] = f.xest[0]; o[1] = f.xest[1];
        o[2] = f.pest[0]; o[3] = f.pest[1]; o[4] = f.pest[2]; o[5] = f.pest[3];
        o[6] = f.nu; o[7] = f.sob;
    }
    free(est);
    return 0;
}

/* g*u[count] as ProcessModel::Predict adds it (system_models.hpp:46-56,82): out[n][2]. */
void ko_gu_table(double dt, int n, double *out)
{
    ko_model *m = malloc(sizeof *m);
    ko_model_init(m, 1.0, 1.0, dt);
    for (int i = 0; i < n && i < KO_U_TABLE; ++i) { out[2 * i] = m->g[0] * m->u[i]; out[2 * i + 1] = m->g[1] * m->u[i]; }
    free(m);
}

/* Data-independent Riccati trace for known-answer tests: S, K, P after step k
 * (kalman_filter.cpp:23-37 only).  out[k] = {S, K0, K1, P00, P01, P10, P11}. */
void ko_riccati_trace(double dt, double q_est, double r_est, int T, double *out)
{
    ko_model *est = malloc(sizeof *est);
    ko_model_init(est, q_est, r_est, dt);
    ko_kf f;
    f.xest[0] = f.xest[1] = 0; f.pest[0] = 1; f.pest[1] = 0; f.pest[2] = 0; f.pest[3] = 1;
    for (int k = 0; k < T; ++k) {
        double P[4]; memcpy(P, f.pest, sizeof P);
        kf_step(est, &f, 0.0, k);
        /* kf_step keeps the gain local; recompute it here with the same operations */
        const double *F = est->Fd;
        double FP0 = F[0] * P[0] + F[1] * P[2], FP1 = F[0] * P[1] + F[1] * P[3];
        double FP2 = F[2] * P[0] + F[3] * P[2], FP3 = F[2] * P[1] + F[3] * P[3];
        double pp0 = (FP0 * F[0] + FP1 * F[1]) + est->Qd[0], pp1 = (FP0 * F[2] + FP1 * F[3]) + est->Qd[1];
        double pp2 = (FP2 * F[0] + FP3 * F[1]) + est->Qd[2], pp3 = (FP2 * F[2] + FP3 * F[3]) + est->Qd[3];
        double c0 = pp0 * est->H[0] + pp1 * est->H[1], c1 = pp2 * est->H[0] + pp3 * est->H[1];
        double sinv = 1.0 / f.sob;
        out[7 * k + 0] = f.sob; out[7 * k + 1] = c0 * sinv; out[7 * k + 2] = c1 * sinv;
        out[7 * k + 3] = f.pest[0]; out[7 * k + 4] = f.pest[1];
        out[7 * k + 5] = f.pest[2]; out[7 * k + 6] = f.pest[3];
    }
    free(est);
}
