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

int aug_tiles(size_t ntrials) { return (int)((ntrials * kAugGroup + kThreads - 1) / kThreads); }

cudaError_t launch_rollout_aug(int N, bool supplied, const AugCase *d_cases, int ncases, int nk, int k0, const AugSampleTime *d_sts,
                               const double *d_u, const void *d_logtab, uint64_t seed, uint32_t trial0, uint32_t ntrials,
                               const double *d_x0noise, const double *d_w, const double *d_v, size_t B, double *d_per_trial,
                               size_t per_trial_ld, double *d_partials, unsigned *d_tickets, double *d_sums, cudaStream_t stream)
{
    if (ncases <= 0 || ntrials == 0) return cudaSuccess;
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
