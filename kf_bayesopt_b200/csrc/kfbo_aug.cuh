// kfbo_aug.cuh -- the STATE-AUGMENTED variants of the rollout (state dimension N = 3, 4), included by kfbo_kernels.cu.
//
// BASELINE.json's north_star: "one trajectory per thread (or per warp for the state-augmented variants)".  The reference has
// no such variant (include/system_models.hpp:11-12 fixes N = 2, Me = 1): what is built here is the reference's recursion
// on a longer chain of integrators, spelled out in DESIGN.md section 4b and pinned by the tests' CPU checker (position, velocity, acceleration
// [, jerk]; white noise and the control drive the highest derivative; H = [1 0 ..]; R = onoise/dt).
//
// Layout: a trajectory is spread over the lanes of a warp, ONE COVARIANCE ROW PER LANE -- lane i of a trajectory's lane
// group holds x_i, xhat_i and row i of P.  A group is kAugGroup = 4 lanes (the smallest power of two that holds N <= 4
// rows), so a warp carries 8 trajectories side by side; for N up to 32 the same code with a 32-lane group is one
// trajectory per warp.  What crosses lanes goes through warp shuffles (width = the group):
//   F x, F xhat              N + N shuffles  (every lane needs the whole state)
//   Ppred = F (P F^T) + Q    P F^T is lane-local (F is upper-triangular Toeplitz: F(j,k) = dt^(k-j)/(k-j)!, uniform
//                            operands); the product with F from the left needs rows i+1.. of it: N (N-1) shuffles
//   S, K, nu                 2 shuffles (c_0 and xpred_0 live in lane 0)
//   P = (I - K H) Ppred      row 0 of Ppred to everyone: N shuffles.  Row 0 itself is formed as R (Ppred_0j / S) -- the same
//                            number as (1 - K_0) Ppred_0j without the cancellation when K_0 -> 1
//   NEES = e^T P^-1 e        forward elimination of [P | e] across the lanes (P = L D L^T, y = L^-1 e, NEES = sum y_k^2 / d_k):
//                            row k to the lanes below it, N (N+1)/2 + N shuffles; no back-substitution, no inverse
// Per lane and step (N = 4): ~100 FP64 instructions and ~45 double shuffles for a quarter of a trajectory -- against the
// per-thread kernel's 48 for a whole N = 2 trajectory: the price of rows across lanes is that the scalar part of the
// step (S, 1/S, nu, NIS, the sums) is replicated in every lane.  What it buys is register space: a thread-per-trajectory
// kernel carries N (N+1)/2 covariance entries, the noise factor and Qd per thread, which stops fitting beyond N ~ 5.
//
// Noise (Philox mode; restated draw for draw by the tests' CPU checker): counter = (trial, block, stream, i ^ stream_x),
// block 0: lane i's x0 draw; block m+1: the step pair 2m, 2m+1 -- words (x, y) -> lane i's process normals of the two
// steps, words (z, w) of lane 0 -> the two observation normals.  w = L n with L's row i in lane i (N shuffles).
#pragma once

namespace kfbo {

__device__ __forceinline__ double aug_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src, kAugGroup); }
__device__ __forceinline__ double aug_shfl_down(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d, kAugGroup); }

template <int N>
struct AugLane {
    double F[N];        // row i of Fd (zero for the padding lanes i >= N)
    double Fdn[N];      // Fdn[d] = Fd(i, i+d) if i + d < N else 0: the coefficient of the row d lanes below
    double L[N];        // row i of the truth noise factor
    double Q[N];        // row i of the estimator's Qd
    double fpow[N];     // Fd(0, j) = dt^j / j!  (uniform: Fd(j, k) = fpow[k - j])
    double g, sr, r_est;
};

// One filter-step of lane i.  nproc: this lane's process-noise input (standard normal in Philox mode, the post-transform
// draw w_i in supplied mode); vobs: the observation noise as lane 0 sees it (standard normal / post-transform draw).
template <int N, bool kSupplied, bool kFirst>
__device__ __forceinline__ void aug_step(const AugLane<N> &m, const int i, double &x, double &xh, double (&P)[N], const double u,
                                         const double nproc, const double vobs, double &kn, double &ki, double &sdn, double &sqn,
                                         double &sdi, double &sqi)
{
    // ---- truth: x = Fd x + g u + w ; z = H x + v          (simulator.cpp:15-19, system_models.hpp:80-100,152-162)
    const double gu = m.g * u;
    double xn = gu, xp = gu;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        xn = fma(m.F[j], aug_shfl(x, j), xn);
        xp = fma(m.F[j], aug_shfl(xh, j), xp);                // predict: xpred = Fd xest + g u  (kalman_filter.cpp:21-22)
    }
    if (kSupplied) {
        xn += nproc;
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) xn = fma(m.L[j], aug_shfl(nproc, j), xn);     // w = L n, row i of L here
    }
    x = xn;
    const double zl = kSupplied ? x + vobs : fma(m.sr, vobs, x);
    const double z = aug_shfl(zl, 0);
    // ---- Ppred = Fd P Fd^T + Qd                            (kalman_filter.cpp:23)
    double B[N], PP[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {                             // B = P Fd^T, row i: B_ij = sum_{k >= j} P_ik Fd(j,k)
        double b = P[j];
#pragma unroll
        for (int k = j + 1; k < N; ++k) b = fma(P[k], m.fpow[k - j], b);
        B[j] = b;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {                             // Ppred = Fd B + Qd, row i: B_ij + sum_d Fd(i,i+d) B_(i+d)j
        double p = B[j] + m.Q[j];
#pragma unroll
        for (int d = 1; d < N; ++d) p = fma(m.Fdn[d], aug_shfl_down(B[j], d), p);
        PP[j] = p;
    }
    // ---- gain: c = Ppred H^T (column 0), S = c_0 + R, K = c (1/S)   (kalman_filter.cpp:26-29)
    const double c = PP[0];
    const double s = aug_shfl(c, 0) + m.r_est;
    const double sinv = KFBO_RCP(s);
    const double kgain = c * sinv;
    // ---- update: nu = z - H xpred ; xest = xpred + K nu ; P = (I - K H) Ppred   (:31-37)
    const double nu = z - aug_shfl(xp, 0);
    xh = fma(kgain, nu, xp);
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double t = aug_shfl(PP[j], 0) * sinv;           // Ppred_0j / S
        const double general = fma(-c, t, PP[j]);             // Ppred_ij - K_i Ppred_0j
        P[j] = (i == 0) ? m.r_est * t : general;              // row 0: (1 - K_0) Ppred_0j = R Ppred_0j / S, no cancellation
    }
    // ---- NEES = e^T P^-1 e by forward elimination across the lanes (trial.cpp:53-58)
    double M[N], y = x - xh, nees = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) M[j] = P[j];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double dk = aug_shfl(M[k], k), yk = aug_shfl(y, k);
        const double rinv = KFBO_RCP(dk);
        nees = fma(yk * yk, rinv, nees);
        if (k + 1 < N) {
            const double f = (i > k) ? M[k] * rinv : 0.0;     // rows above and at the pivot stay as they are
#pragma unroll
            for (int j = k + 1; j < N; ++j) M[j] = fma(-f, aug_shfl(M[j], k), M[j]);
            y = fma(-f, yk, y);
        }
    }
    const double nu_s = nu * sinv;                            // NIS = (nu / S) nu  (trial.cpp:59-60)
    if (kFirst) { kn = nees; ki = nu_s * nu; }                // shift of the one-pass variance sums: the first sample
    const double dn = nees - kn, di = fma(nu_s, nu, -ki);
    sdn += dn; sqn = fma(dn, dn, sqn);
    sdi += di; sqi = fma(di, di, sqi);
}

template <int N, bool kSupplied>
__global__ void __launch_bounds__(kThreads, kSupplied ? 1 : 4)
rollout_aug(const AugCase *__restrict__ cases, const AugSampleTime *__restrict__ sts, int k0, const double *__restrict__ u_tables,
            const math::LogTables *__restrict__ logtab, const __grid_constant__ PhiloxKeys key, uint32_t trial0, uint32_t ntrials,
            int tiles, const double *__restrict__ x0noise, const double *__restrict__ w, const double *__restrict__ v, size_t B,
            double *__restrict__ per_trial, size_t per_trial_ld, double *__restrict__ partials, unsigned *__restrict__ tickets,
            double *__restrict__ sums)
{
    static_assert(N >= 2 && N <= kAugGroup, "one covariance row per lane of a kAugGroup-lane group");
    extern __shared__ __align__(16) unsigned char s_tables_raw[];
    math::LogTables &s_log = *reinterpret_cast<math::LogTables *>(s_tables_raw);
    const int case_idx = blockIdx.y * gridDim.z + blockIdx.z, tile = blockIdx.x;
    const AugCase &cs = cases[case_idx];
    const AugSampleTime &st = sts[k0 + blockIdx.z];
    const double *__restrict__ u_tab = u_tables + (size_t)(k0 + blockIdx.z) * kMaxSteps;
    const int T = cs.T;
    const int i = threadIdx.x & (kAugGroup - 1);
    const uint32_t local = ((uint32_t)tile * kThreads + threadIdx.x) / kAugGroup;     // trajectory of this launch
    const bool live = local < ntrials;
    const uint32_t trial = trial0 + (live ? local : ntrials - 1u);                    // padding groups repeat the last one
    AugLane<N> m;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const bool row = i < N;
        m.F[j] = row ? st.F[i * kAugGroup + j] : 0.0;
        m.L[j] = row ? st.L[i * kAugGroup + j] : 0.0;
        m.Q[j] = row ? cs.qd[i * kAugGroup + j] : 0.0;
        m.Fdn[j] = (i + j < N) ? st.F[i * kAugGroup + i + j] : 0.0;
        m.fpow[j] = st.F[j];
    }
    m.g = i < N ? st.g[i] : 0.0;
    m.sr = st.sr;
    m.r_est = cs.r_est;
    if (!kSupplied) {
        load_noise_tables(&s_log, logtab);
        __syncthreads();
    }
    double x, xh = 0.0, P[N];                                                         // x0 = 0, P0 = I (trial.cpp:31-32)
#pragma unroll
    for (int j = 0; j < N; ++j) P[j] = (i == j) ? 1.0 : 0.0;
    double kn = 0.0, ki = 0.0, sdn = 0.0, sqn = 0.0, sdi = 0.0, sqi = 0.0;
    if (kSupplied) {
        const size_t col = (size_t)(trial - trial0);
        x = i < N ? x0noise[(size_t)i * B + col] : 0.0;                               // x = x0 + N(0, P0) (simulator.cpp:12)
        const double *pw = w + (size_t)(i < N ? i : 0) * B + col, *pv = v + col;
        for (int s = 0; s < T; ++s) {
            const double wi = i < N ? __ldg(pw + (size_t)s * N * B) : 0.0, vv = __ldg(pv + (size_t)s * B);
            if (s == 0) aug_step<N, true, true>(m, i, x, xh, P, __ldg(u_tab + s), wi, vv, kn, ki, sdn, sqn, sdi, sqi);
            else        aug_step<N, true, false>(m, i, x, xh, P, __ldg(u_tab + s), wi, vv, kn, ki, sdn, sqn, sdi, sqi);
        }
    } else {
        {
            double a, b;
            const uint4 r0 = philox_at(trial, cs.stream, cs.stream_x, 0u, (uint32_t)i, key);
            math::normal_pair(r0.x, r0.y, s_log, a, b);
            x = a;
        }
        const int npairs = (T + 1) >> 1;
        uint4 r = philox_at(trial, cs.stream, cs.stream_x, 1u, (uint32_t)i, key);
        for (int p = 0; p < npairs; ++p) {
            const uint4 q = philox_at(trial, cs.stream, cs.stream_x, (uint32_t)p + 2u, (uint32_t)i, key);   // the next pair's bits, early
            double na, nb, oa, ob;
            math::normal_pair(r.x, r.y, s_log, na, nb);
            math::normal_pair(r.z, r.w, s_log, oa, ob);
            const int s = 2 * p;
            if (p == 0) aug_step<N, false, true>(m, i, x, xh, P, __ldg(u_tab + s), na, oa, kn, ki, sdn, sqn, sdi, sqi);
            else        aug_step<N, false, false>(m, i, x, xh, P, __ldg(u_tab + s), na, oa, kn, ki, sdn, sqn, sdi, sqi);
            if (s + 1 < T) aug_step<N, false, false>(m, i, x, xh, P, __ldg(u_tab + s + 1), nb, ob, kn, ki, sdn, sqn, sdi, sqi);
            r = q;
        }
    }
    double out[4] = {0.0, 0.0, 0.0, 0.0};
    {
        Traj t;
        t.kn = kn; t.ki = ki; t.sdn = sdn; t.sqn = sqn; t.sdi = sdi; t.sqi = sqi;
        double o[4];
        traj_finish(t, T, o);
        if (live && i == 0) {                                                         // one lane of the group reports the trajectory
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k] = o[k];
            if (per_trial != nullptr) {
#pragma unroll
                for (int k = 0; k < 4; ++k) per_trial[((size_t)cs.out_slot * 4 + k) * per_trial_ld + local] = o[k];
            }
        }
    }
    const int out_slot = cs.out_slot;
    __syncthreads();
    __shared__ double s_red_static[kSupplied ? kThreads : 1][4];
    block_reduce_and_publish(out, case_idx, tile, tiles, out_slot, partials, tickets, sums,
                             kSupplied ? s_red_static : reinterpret_cast<double (*)[4]>(s_tables_raw));
}

// -----------------------------------------------------------------------------------------
// The same variants with ONE TRAJECTORY PER THREAD (KFBO_AUG_LAYOUT_THREAD, the default for on-device noise): x, xhat and
// the whole N x N covariance in the thread's registers, nothing crosses lanes.  It performs lane i's arithmetic of
// aug_step above for every row i, operation for operation (the products with the structural zeros of Fd and L, which are
// exact no-ops there, are left out), and draws the same Philox words for the same (trial, lane, step) -- so every
// trajectory's outputs are the SAME BITS as the rows-across-lanes kernel's (the cost sums add them in another order;
// test_gpu_aug.py), and the tests' CPU checker pins both.  Why it
// exists: rows across lanes replicate the scalar part of the step in every lane and pay 45 double shuffles per lane-step;
// per filter-step this kernel issues under a third of the instructions (N = 4: 234 FP64 + 2 Philox calls against 450
// FP64 + 350 SHFL).  Why the other one stays: registers -- N (N + 2) doubles of state per thread stop fitting
// beyond N = 4, the lane layout scales to N = 32 -- and it is the layout BASELINE.json's north_star names.
// The sample time's constants (Fd, g, L, sqrt R) come by value: one launch per sample time, every use is a
// constant-bank operand of the FP64 instruction (no register).  The case's Qd sits in shared memory.
// -----------------------------------------------------------------------------------------
// CTAs per SM the per-thread kernels are compiled for (registers per thread <= 65536 / (128 x this)).  Measured
// (profiles/variants_aug_r2q.txt): N = 3 fits 126 registers without a spill, 4 CTAs per SM 6.40 ms against 6.63 with 3;
// N = 4 needs 168 (3 CTAs: 9.86 ms; 2 CTAs with 202 registers and no spill at all: 10.13).
#ifndef KFBO_AUG_THREAD_MIN_BLOCKS
#define KFBO_AUG_THREAD_MIN_BLOCKS(N) ((N) == 3 ? 4 : 3)
#endif
template <int N>
struct AugThread {
    double x[N], xh[N], P[N][N];
    double kn, ki, sdn, sqn, sdi, sqi;
};

template <int N, bool kFirst>
__device__ __forceinline__ void aug_thread_step(const AugSampleTime &st, const double *__restrict__ q /* [N][kAugGroup], shared */,
                                                const double r_est, AugThread<N> &t, const double u, const double (&nproc)[N],
                                                const double vobs)
{
    // ---- truth and predict: x = Fd x + g u + L n ; xpred = Fd xest + g u ; z = x_0 + sqrt(R) v
    double xn[N], xp[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double gu = st.g[i] * u;
        double a = gu, b = gu;
#pragma unroll
        for (int j = i; j < N; ++j) {
            a = fma(st.F[i * kAugGroup + j], t.x[j], a);
            b = fma(st.F[i * kAugGroup + j], t.xh[j], b);
        }
#pragma unroll
        for (int j = 0; j <= i; ++j) a = fma(st.L[i * kAugGroup + j], nproc[j], a);
        xn[i] = a; xp[i] = b;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) t.x[i] = xn[i];
    const double z = fma(st.sr, vobs, t.x[0]);
    // ---- Ppred = Fd (P Fd^T) + Qd ; Fd(j, k) = Fd(0, k - j)
    double B[N][N], PP[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double b = t.P[i][j];
#pragma unroll
            for (int k = j + 1; k < N; ++k) b = fma(t.P[i][k], st.F[k - j], b);
            B[i][j] = b;
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double p = B[i][j] + q[i * kAugGroup + j];
#pragma unroll
            for (int d = 1; i + d < N; ++d) p = fma(st.F[i * kAugGroup + i + d], B[i + d][j], p);
            PP[i][j] = p;
        }
    }
    // ---- gain, update
    const double s = PP[0][0] + r_est;
    const double sinv = KFBO_RCP(s);
    const double nu = z - xp[0];
    double tj[N];
#pragma unroll
    for (int j = 0; j < N; ++j) tj[j] = PP[0][j] * sinv;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double c = PP[i][0];
        t.xh[i] = fma(c * sinv, nu, xp[i]);
#pragma unroll
        for (int j = 0; j < N; ++j) t.P[i][j] = (i == 0) ? r_est * tj[j] : fma(-c, tj[j], PP[i][j]);
    }
    // ---- NEES by forward elimination of [P | e]
    double M[N][N], y[N], nees = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        y[i] = t.x[i] - t.xh[i];
#pragma unroll
        for (int j = 0; j < N; ++j) M[i][j] = t.P[i][j];
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double rinv = KFBO_RCP(M[k][k]);
        nees = fma(y[k] * y[k], rinv, nees);
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const double f = M[i][k] * rinv;
#pragma unroll
            for (int j = k + 1; j < N; ++j) M[i][j] = fma(-f, M[k][j], M[i][j]);
            y[i] = fma(-f, y[k], y[i]);
        }
    }
    const double nu_s = nu * sinv;
    if (kFirst) { t.kn = nees; t.ki = nu_s * nu; }
    const double dn = nees - t.kn, di = fma(nu_s, nu, -t.ki);
    t.sdn += dn; t.sqn = fma(dn, dn, t.sqn);
    t.sdi += di; t.sqi = fma(di, di, t.sqi);
}

template <int N>
__global__ void __launch_bounds__(kThreads, KFBO_AUG_THREAD_MIN_BLOCKS(N))
rollout_aug_thread(const AugCase *__restrict__ cases, const __grid_constant__ AugSampleTime st, int nk, int k,
                   const double *__restrict__ u_tab, const math::LogTables *__restrict__ logtab, const __grid_constant__ PhiloxKeys key,
                   uint32_t trial0, uint32_t ntrials, int tiles, double *__restrict__ per_trial, size_t per_trial_ld,
                   double *__restrict__ partials, unsigned *__restrict__ tickets, double *__restrict__ sums)
{
    static_assert(N >= 2 && N <= kAugGroup, "matrices are stored with a row stride of kAugGroup");
    extern __shared__ __align__(16) unsigned char s_tables_raw[];
    __shared__ double s_q[N * kAugGroup];
    math::LogTables &s_log = *reinterpret_cast<math::LogTables *>(s_tables_raw);
    const int case_idx = blockIdx.y * nk + k, tile = blockIdx.x;
    const AugCase &cs = cases[case_idx];
    const int T = cs.T;
    const double r_est = cs.r_est;
    const uint32_t stream = cs.stream, stream_x = cs.stream_x;
    const int out_slot = cs.out_slot;
    if (threadIdx.x < N * kAugGroup) s_q[threadIdx.x] = cs.qd[threadIdx.x];
    load_noise_tables(&s_log, logtab);
    __syncthreads();
    const uint32_t local = (uint32_t)tile * kThreads + threadIdx.x;
    const bool live = local < ntrials;
    const uint32_t trial = trial0 + (live ? local : ntrials - 1u);   // padding threads repeat the last trajectory
    AugThread<N> t;
#pragma unroll
    for (int i = 0; i < N; ++i) {                                    // x = x0 + N(0, P0), x0 = 0, P0 = I; xhat = 0
        double a, b;
        const uint4 r0 = philox_at(trial, stream, stream_x, 0u, (uint32_t)i, key);
        math::normal_pair(r0.x, r0.y, s_log, a, b);
        t.x[i] = a; t.xh[i] = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) t.P[i][j] = (i == j) ? 1.0 : 0.0;
    }
    t.kn = 0.0; t.ki = 0.0; t.sdn = 0.0; t.sqn = 0.0; t.sdi = 0.0; t.sqi = 0.0;
    const int npairs = (T + 1) >> 1;
    for (int p = 0; p < npairs; ++p) {
        // the step pair 2p, 2p+1: call i's words (x, y) -> row i's process normals of the two steps, call 0's (z, w) -> the observation normals
        double na[N], nb[N], oa, ob;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint4 r = philox_at(trial, stream, stream_x, (uint32_t)p + 1u, (uint32_t)i, key);
            math::normal_pair(r.x, r.y, s_log, na[i], nb[i]);
            if (i == 0) math::normal_pair(r.z, r.w, s_log, oa, ob);
        }
        const int s = 2 * p;
        if (p == 0) aug_thread_step<N, true>(st, s_q, r_est, t, __ldg(u_tab + s), na, oa);
        else        aug_thread_step<N, false>(st, s_q, r_est, t, __ldg(u_tab + s), na, oa);
        if (s + 1 < T) aug_thread_step<N, false>(st, s_q, r_est, t, __ldg(u_tab + s + 1), nb, ob);
    }
    double out[4] = {0.0, 0.0, 0.0, 0.0};
    {
        Traj tf;
        tf.kn = t.kn; tf.ki = t.ki; tf.sdn = t.sdn; tf.sqn = t.sqn; tf.sdi = t.sdi; tf.sqi = t.sqi;
        double o[4];
        traj_finish(tf, T, o);
        if (live) {
#pragma unroll
            for (int m = 0; m < 4; ++m) out[m] = o[m];
            if (per_trial != nullptr) {
#pragma unroll
                for (int m = 0; m < 4; ++m) per_trial[((size_t)out_slot * 4 + m) * per_trial_ld + local] = o[m];
            }
        }
    }
    __syncthreads();
    block_reduce_and_publish(out, case_idx, tile, tiles, out_slot, partials, tickets, sums, reinterpret_cast<double (*)[4]>(s_tables_raw));
}

int aug_tiles(size_t ntrials) { return (int)((ntrials * kAugGroup + kThreads - 1) / kThreads); }

cudaError_t launch_rollout_aug(int N, bool supplied, const AugCase *d_cases, int ncases, int nk, int k0, const AugSampleTime *d_sts,
                               const double *d_u, const void *d_logtab, uint64_t seed, uint32_t trial0, uint32_t ntrials,
                               const double *d_x0noise, const double *d_w, const double *d_v, size_t B, double *d_per_trial,
                               size_t per_trial_ld, double *d_partials, unsigned *d_tickets, double *d_sums, cudaStream_t stream,
                               const AugSampleTime *h_sts, int layout)
{
    if (ncases <= 0 || ntrials == 0) return cudaSuccess;
    if (!supplied && layout == kAugLayoutThread && h_sts != nullptr) {
        // one trajectory per thread: one launch per sample time, its constants by value
        const int tiles = rollout_tiles(ntrials);
        const dim3 grid((unsigned)tiles, (unsigned)(ncases / nk), 1u);
        const size_t smem = sizeof(math::LogTables);
        auto go = [&](auto *kernel) -> cudaError_t {
            const cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(kernel), 0, smem);
            if (e != cudaSuccess) return e;
            for (int k = 0; k < nk; ++k) {
                kernel<<<grid, kThreads, smem, stream>>>(d_cases, h_sts[k0 + k], nk, k, d_u + (size_t)(k0 + k) * kMaxSteps,
                                                          static_cast<const math::LogTables *>(d_logtab), expand_philox_key(seed), trial0,
                                                          ntrials, tiles, d_per_trial, per_trial_ld, d_partials, d_tickets, d_sums);
                const cudaError_t el = cudaGetLastError();
                if (el != cudaSuccess) return el;
            }
            return cudaSuccess;
        };
        if (N == 3) return go(rollout_aug_thread<3>);
        if (N == 4) return go(rollout_aug_thread<4>);
        return cudaErrorInvalidValue;
    }
    const int tiles = aug_tiles(ntrials);
    const dim3 grid((unsigned)tiles, (unsigned)(ncases / nk), (unsigned)nk);
    const size_t smem = supplied ? 0 : sizeof(math::LogTables);
    auto go = [&](auto *kernel, int slot) -> cudaError_t {
        if (smem) {
            const cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(kernel), slot, smem);
            if (e != cudaSuccess) return e;
        }
        kernel<<<grid, kThreads, smem, stream>>>(d_cases, d_sts, k0, d_u, static_cast<const math::LogTables *>(d_logtab),
                                                  expand_philox_key(seed), trial0, ntrials, tiles, d_x0noise, d_w, d_v, B,
                                                  d_per_trial, per_trial_ld, d_partials, d_tickets, d_sums);
        return cudaGetLastError();
    };
    if (N == 3) return supplied ? go(rollout_aug<3, true>, 10) : go(rollout_aug<3, false>, 11);
    if (N == 4) return supplied ? go(rollout_aug<4, true>, 12) : go(rollout_aug<4, false>, 13);
    return cudaErrorInvalidValue;
}

}  // namespace kfbo
