/*
 * kfbo_oracle_aug.c -- CPU ORACLE for the STATE-AUGMENTED variants (state dimension N = 3, 4) of the Monte Carlo
 * cost evaluation.  TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as kfbo_oracle.c: only tests/, smoke() and
 * bench.py's checker legs may load it).
 *
 * PARITY UNPINNED, and necessarily so: the reference has NO such variant -- N and Me are #defines fixed at 2 and 1
 * (include/system_models.hpp:11-12) and only pnoise_[0] / onoise_[0] are ever read (:39,143).  BASELINE.json's north_star
 * names "state-augmented variants" all the same, so this file states what the reference's formulas become when its
 * state is a longer chain of integrators, changing nothing else:
 *
 *   Fc = ones on the superdiagonal (system_models.hpp:35-37 for N = 2: [[0,1],[0,0]]), Qc(N-1,N-1) = pnoise_[0] (:39)
 *   g_i = dt^(N-i)/(N-i)!   (:46-47 for N = 2: [dt^2/2, dt] -- the control drives the highest derivative, like the noise)
 *   u_[i] = 2cos(0.75 j), j += dt (:49-56);  H = [1 0 ... 0], R = onoise_[0]/dt (:141-143)
 *   continuousToDiscrete by Van Loan on the 2N x 2N matrix (include/continuous2discrete.hpp:8-32)
 *   x0 = 0, P0 = I (trial.cpp:31-32);  truth / filter / NEES / NIS / time variance exactly as trial.cpp:45-98,
 *   simulator.cpp:15-19, kalman_filter.cpp:18-38 with N x N products in coefficient order and P^-1 by cofactors x 1/det
 *   (what Eigen's fixed-size inverse does up to 4 x 4).
 *
 * What pins it instead: scipy.linalg.expm and a numpy restatement of the filter (tests/test_oracle_aug.py), the
 * closed forms Fd_ij = dt^(j-i)/(j-i)!, Qd_ij = q dt^(2N-1-i-j) / ((N-1-i)! (N-1-j)! (2N-1-i-j)), chi-square consistency
 * (E[NEES] = N, E[NIS] = 1 at matched parameters), and -- for N = 2 -- bit-for-bit agreement with kfbo_oracle.c.
 *
 * Compile with -O2 -ffp-contract=off like kfbo_oracle.c.  Matrices are row-major double[N*N], N <= KOA_MAXN.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KOA_MAXN 4
#define KOA_U_TABLE 2000

void ko_control_table(double dt, int n, double *u);
void ko_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out);
void ko_normal_pair(uint32_t w0, uint32_t w1, double *na, double *nb);

/* exp(A) for an n x n matrix, n <= 8: Taylor series with scaling and squaring (the same definition kfbo_oracle.c's
 * ko_expm4 restates for Eigen's MatrixBase::exp(), continuous2discrete.hpp:25).  The Van Loan matrices of the
 * integrator chains are nilpotent (A^(2N) = 0), so the series terminates. */
static void matn_mul(int n, const double *a, const double *b, double *c)
{
    double t[64];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += a[n * i + k] * b[n * k + j];
            t[n * i + j] = s;
        }
    memcpy(c, t, sizeof(double) * n * n);
}

void koa_expm(int n, const double *A, double *E)
{
    double norm1 = 0.0;
    for (int j = 0; j < n; ++j) {
        double cs = 0.0;
        for (int i = 0; i < n; ++i) cs += fabs(A[n * i + j]);
        if (cs > norm1) norm1 = cs;
    }
    int s = 0;
    double scale = 1.0;
    while (norm1 * scale > 0.5 && s < 60) { scale *= 0.5; ++s; }
    double As[64], term[64], sum[64];
    for (int i = 0; i < n * n; ++i) As[i] = A[i] * scale;
    for (int i = 0; i < n * n; ++i) sum[i] = term[i] = (i % (n + 1) == 0) ? 1.0 : 0.0;
    for (int k = 1; k <= 60; ++k) {
        matn_mul(n, term, As, term);
        double mx = 0.0;
        for (int i = 0; i < n * n; ++i) { term[i] /= (double)k; if (fabs(term[i]) > mx) mx = fabs(term[i]); }
        for (int i = 0; i < n * n; ++i) sum[i] += term[i];
        if (mx == 0.0) break;
    }
    for (int i = 0; i < s; ++i) matn_mul(n, sum, sum, sum);
    memcpy(E, sum, sizeof(double) * n * n);
}

/* continuous2discrete.hpp:8-32 for the N-integrator chain: bigA = [[-Fc dt, Qc dt],[0, Fc^T dt]], bigB = exp(bigA),
 * Fd = bottomRight^T, Qd = Fd * topRight. */
int koa_discretize(int N, double q, double dt, double *Fd, double *Qd)
{
    if (N < 2 || N > KOA_MAXN) return -1;
    const int n = 2 * N;
    double bigA[64], bigB[64];
    memset(bigA, 0, sizeof bigA);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            const double fc = (j == i + 1) ? 1.0 : 0.0, qc = (i == N - 1 && j == N - 1) ? q : 0.0;
            const double fct = (i == j + 1) ? 1.0 : 0.0;
            bigA[n * i + j] = -fc * dt;
            bigA[n * i + (j + N)] = qc * dt;
            bigA[n * (i + N) + (j + N)] = fct * dt;
        }
    koa_expm(n, bigA, bigB);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) Fd[N * i + j] = bigB[n * (j + N) + (i + N)];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
            for (int k = 0; k < N; ++k) s += Fd[N * i + k] * bigB[n * k + (j + N)];
            Qd[N * i + j] = s;
        }
    return 0;
}

typedef struct {
    int N;
    double Fd[16], Qd[16], g[4], R;
    double u[KOA_U_TABLE];
} koa_model;

static void koa_model_init(koa_model *m, int N, double pnoise0, double onoise0, double dt)
{
    m->N = N;
    for (int i = 0; i < N; ++i) {             /* g_i = dt^(N-i)/(N-i)!, built like 0.5*dt*dt / dt of system_models.hpp:46-47 */
        double p = 1.0, f = 1.0;
        for (int k = 1; k <= N - i; ++k) { p *= dt; f *= (double)k; }
        m->g[i] = p / f;
    }
    ko_control_table(dt, KOA_U_TABLE, m->u);
    koa_discretize(N, pnoise0, dt, m->Fd, m->Qd);
    m->R = onoise0 / dt;
}

void koa_gvector(int N, double dt, double *g)
{
    koa_model *m = malloc(sizeof *m);
    koa_model_init(m, N, 1.0, 1.0, dt);
    memcpy(g, m->g, sizeof(double) * N);
    free(m);
}

/* xpred = fd_*x + g_*u_[count] (system_models.hpp:82), coefficient-order products */
static void process_predict(const koa_model *m, const double *x, int count, double *xp)
{
    const int N = m->N;
    const double u = m->u[count];
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        for (int k = 0; k < N; ++k) s += m->Fd[N * i + k] * x[k];
        xp[i] = s + m->g[i] * u;
    }
}

/* determinant and cofactor inverse of an N x N matrix, N <= 4 (minors expanded recursively along the first row) */
static double detn(int n, const double *a)
{
    if (n == 1) return a[0];
    if (n == 2) return a[0] * a[3] - a[1] * a[2];
    double d = 0.0, sub[9];
    for (int c = 0; c < n; ++c) {
        int p = 0;
        for (int i = 1; i < n; ++i)
            for (int j = 0; j < n; ++j)
                if (j != c) sub[p++] = a[n * i + j];
        const double t = a[c] * detn(n - 1, sub);
        d += (c & 1) ? -t : t;
    }
    return d;
}

static void inverse_cofactor(int n, const double *a, double *inv)
{
    const double invdet = 1.0 / detn(n, a);
    double sub[9];
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            int p = 0;
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j)
                    if (i != r && j != c) sub[p++] = a[n * i + j];
            const double cof = (((r + c) & 1) ? -1.0 : 1.0) * detn(n - 1, sub);
            inv[n * c + r] = cof * invdet;          /* adjugate = transposed cofactors */
        }
}

typedef struct { double xest[4], pest[16], nu, sob; } koa_kf;

/* KalmanFilter::Step, kalman_filter.cpp:18-38, N x N */
static void kf_step(const koa_model *m, koa_kf *f, double z, int count)
{
    const int N = m->N;
    double xpred[4], FP[16], ppred[16], c[4] = {0.0, 0.0, 0.0, 0.0}, k[4], X[16], P2[16];
    process_predict(m, f->xest, count, xpred);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
            for (int l = 0; l < N; ++l) s += m->Fd[N * i + l] * f->pest[N * l + j];
            FP[N * i + j] = s;
        }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
            for (int l = 0; l < N; ++l) s += FP[N * i + l] * m->Fd[N * j + l];
            ppred[N * i + j] = s + m->Qd[N * i + j];     /* :23 */
        }
    for (int i = 0; i < N; ++i) c[i] = ppred[N * i];     /* c = Ppred H^T, H = [1 0 ...] (:26; products with 0 / 1 are exact) */
    f->sob = c[0] + m->R;                                /* :27 */
    const double sinv = 1.0 / f->sob;                    /* :28 */
    for (int i = 0; i < N; ++i) k[i] = c[i] * sinv;
    f->nu = z - xpred[0];                                /* :31 */
    for (int i = 0; i < N; ++i) f->xest[i] = xpred[i] + k[i] * f->nu;   /* :33 */
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) X[N * i + j] = ((i == j) ? 1.0 : 0.0) - ((j == 0) ? k[i] : 0.0);   /* I - K H (:36) */
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
            for (int l = 0; l < N; ++l) s += X[N * i + l] * ppred[N * l + j];
            P2[N * i + j] = s;                           /* :37, no symmetrisation */
        }
    memcpy(f->pest, P2, sizeof(double) * N * N);
}

/* trial.cpp:53-60: NEES = (e^T P^-1) e, NIS = (nu / S) nu */
static void chi2(int N, const double *x, const koa_kf *f, double *nees, double *nis)
{
    double e[4], inv[16], r[4];
    for (int i = 0; i < N; ++i) e[i] = x[i] - f->xest[i];
    inverse_cofactor(N, f->pest, inv);
    for (int j = 0; j < N; ++j) {
        double s = 0.0;
        for (int i = 0; i < N; ++i) s += e[i] * inv[N * i + j];
        r[j] = s;
    }
    double s = 0.0;
    for (int j = 0; j < N; ++j) s += r[j] * e[j];
    *nees = s;
    *nis = (f->nu * (1.0 / f->sob)) * f->nu;
}

typedef void (*koa_noise_fn)(void *ctx, int step, double *w /*[N]*/, double *v);

static void run_trial(const koa_model *truth, const koa_model *est, int T, const double *x0n, koa_noise_fn noise, void *ctx,
                      double *his_nees, double *his_nis, double *out)
{
    const int N = truth->N;
    double x[4];
    koa_kf f;
    memset(&f, 0, sizeof f);
    for (int i = 0; i < N; ++i) { x[i] = 0.0 + x0n[i]; f.pest[N * i + i] = 1.0; }   /* trial.cpp:31-32, simulator.cpp:12 */
    f.sob = 1.0;
    double row_nees = 0.0, row_nis = 0.0;
    for (int step = 0; step < T; ++step) {
        double w[4], v, xp[4];
        noise(ctx, step, w, &v);
        process_predict(truth, x, step, xp);
        for (int i = 0; i < N; ++i) x[i] = xp[i] + w[i];
        const double z = x[0] + v;
        kf_step(est, &f, z, step);
        double nees, nis;
        chi2(N, x, &f, &nees, &nis);
        row_nees += nees; row_nis += nis;
        his_nees[step] = nees; his_nis[step] = nis;
    }
    const double ave_nees = row_nees / T, ave_nis = row_nis / T;
    double var_nees = 0.0, var_nis = 0.0;
    for (int i = 0; i < T; ++i) {
        var_nees += (his_nees[i] - ave_nees) * (his_nees[i] - ave_nees);
        var_nis += (his_nis[i] - ave_nis) * (his_nis[i] - ave_nis);
    }
    out[0] = row_nees; out[1] = row_nis; out[2] = var_nees / (T - 1); out[3] = var_nis / (T - 1);
}

/* ---- supplied noise: x0noise[N][B], w[T][N][B], v[T][B] -> per_trial[4][B] ---- */
typedef struct { const double *w, *v; long B, i; int N; } supplied_ctx;

static void supplied_noise(void *p, int step, double *w, double *v)
{
    const supplied_ctx *c = (const supplied_ctx *)p;
    for (int k = 0; k < c->N; ++k) w[k] = c->w[((long)step * c->N + k) * c->B + c->i];
    *v = c->v[(long)step * c->B + c->i];
}

int koa_eval_supplied(int N, double dt, int T, double q_est, double r_est, long B, const double *x0noise, const double *w,
                      const double *v, double *per_trial)
{
    if (N < 2 || N > KOA_MAXN || T < 2 || T > KOA_U_TABLE || B < 0) return -1;
    koa_model *truth = malloc(sizeof *truth), *est = malloc(sizeof *est);
    double *hn = malloc(sizeof(double) * T), *hi = malloc(sizeof(double) * T);
    koa_model_init(truth, N, 1.0, 1.0, dt);     /* the truth noise is supplied: its (q, r) never enter */
    koa_model_init(est, N, q_est, r_est, dt);
    for (long i = 0; i < B; ++i) {
        supplied_ctx c = {w, v, B, i, N};
        double x0n[4], out[4];
        for (int k = 0; k < N; ++k) x0n[k] = x0noise[(long)k * B + i];
        run_trial(truth, est, T, x0n, supplied_noise, &c, hn, hi, out);
        for (int k = 0; k < 4; ++k) per_trial[(long)k * B + i] = out[k];
    }
    free(truth); free(est); free(hn); free(hi);
    return 0;
}

/* lower Cholesky factor of an N x N SPD matrix (the truth noise transform; any T with T T^T = Qd is statistically
 * equivalent to eigenmvn.h:126-130's V sqrt(Lambda)) */
int koa_cholesky(int N, const double *A, double *L)
{
    memset(L, 0, sizeof(double) * N * N);
    for (int j = 0; j < N; ++j) {
        double s = A[N * j + j];
        for (int k = 0; k < j; ++k) s -= L[N * j + k] * L[N * j + k];
        if (!(s > 0.0)) return -1;
        L[N * j + j] = sqrt(s);
        for (int i = j + 1; i < N; ++i) {
            double t = A[N * i + j];
            for (int k = 0; k < j; ++k) t -= L[N * i + k] * L[N * j + k];
            L[N * i + j] = t / L[N * j + j];
        }
    }
    return 0;
}

/* ---- Philox noise, the product's on-device layout for the augmented kernel (kfbo_aug.cuh), restated draw for draw:
 *   counter = (trial, block, stream[31:0], component ^ (stream[61:32] << 2)), key = seed
 *   block 0          component i: words (x, y) -> pair; its FIRST normal is x0noise_i (P0 = I)
 *   block m+1        the STEP PAIR 2m, 2m+1; component i: words (x, y) -> normals (a, b): the standard normal n_i of the process
 *                    noise at step 2m is a, at step 2m+1 is b;  component 0's words (z, w) -> (a', b'): the observation normal
 *                    of step 2m / 2m+1 (the (z, w) words of the other components are not used)
 *   w = L n (L = lower Cholesky factor of the truth Qd), v = sqrt(R) n_obs */
typedef struct {
    uint32_t key[2], trial, stream, stream_x;
    int N;
    double L[16], sr;
    int cached_pair;
    double pn[2][4], on[2];
} philox_ctx;

static void philox_noise(void *p, int step, double *w, double *v)
{
    philox_ctx *c = (philox_ctx *)p;
    const int m = step >> 1, s = step & 1;
    if (m != c->cached_pair) {
        for (uint32_t i = 0; i < (uint32_t)c->N; ++i) {
            const uint32_t ctr[4] = {c->trial, (uint32_t)m + 1u, c->stream, i ^ c->stream_x};
            uint32_t o[4];
            ko_philox4x32_10(ctr, c->key, o);
            ko_normal_pair(o[0], o[1], &c->pn[0][i], &c->pn[1][i]);
            if (i == 0) ko_normal_pair(o[2], o[3], &c->on[0], &c->on[1]);
        }
        c->cached_pair = m;
    }
    for (int i = 0; i < c->N; ++i) {
        double t = 0.0;
        for (int j = 0; j <= i; ++j) t += c->L[c->N * i + j] * c->pn[s][j];
        w[i] = t;
    }
    *v = c->sr * c->on[s];
}

int koa_eval_philox(int N, uint64_t seed, uint64_t stream, double dt, int T, double q_truth, double r_truth, double q_est,
                    double r_est, uint32_t trial0, long ntrials, double *per_trial)
{
    if (N < 2 || N > KOA_MAXN || T < 2 || T > KOA_U_TABLE || ntrials < 0) return -1;
    koa_model *truth = malloc(sizeof *truth), *est = malloc(sizeof *est);
    double *hn = malloc(sizeof(double) * T), *hi = malloc(sizeof(double) * T);
    koa_model_init(truth, N, q_truth, r_truth, dt);
    koa_model_init(est, N, q_est, r_est, dt);
    philox_ctx c;
    memset(&c, 0, sizeof c);
    c.key[0] = (uint32_t)seed; c.key[1] = (uint32_t)(seed >> 32);
    c.stream = (uint32_t)stream;
    c.stream_x = (uint32_t)((stream >> 32) & 0x3FFFFFFFu) << 2;
    c.N = N;
    int rc = koa_cholesky(N, truth->Qd, c.L);
    c.sr = sqrt(truth->R);
    for (long t = 0; t < ntrials && rc == 0; ++t) {
        c.trial = trial0 + (uint32_t)t;
        c.cached_pair = -1;
        double x0n[4], out[4], unused;
        for (uint32_t i = 0; i < (uint32_t)N; ++i) {
            const uint32_t ctr[4] = {c.trial, 0u, c.stream, i ^ c.stream_x};
            uint32_t o[4];
            ko_philox4x32_10(ctr, c.key, o);
            ko_normal_pair(o[0], o[1], &x0n[i], &unused);
        }
        run_trial(truth, est, T, x0n, philox_noise, &c, hn, hi, out);
        for (int k = 0; k < 4; ++k) per_trial[(long)k * ntrials + t] = out[k];
    }
    free(truth); free(est); free(hn); free(hi);
    return rc;
}
