// lone_warp_filter.cu -- how long does ONE warp take per filter_step when nothing else runs on its scheduler?  (The latency floor of
// small launches: rollout_philox_small's consumer warp.)  Variants: the full step; without the NEES / NIS part; covariance recursion only.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../kf_bayesopt_b200/csrc/kfbo_device.cuh"

using namespace kfbo;

template <int MODE>
__global__ void __launch_bounds__(128) k(const Case pc, double *out, int steps, long long *cyc)
{
    const Case m = pc;
    Traj t;
    traj_init<false>(t, 0.01 * threadIdx.x, 0.02);
    double w = 1e-3 * threadIdx.x;
    const long long c0 = clock64();
    for (int s = 0; s < steps; ++s) {
        const double2 g = make_double2(1e-3 * (s & 7), 1e-2 * (s & 3));
        if (MODE == 0) {
            filter_step<false, false, false>(t, m, m.f, g, StepNoise{w, -w, 0.5 * w});
        } else {   // covariance recursion only (the loop-carried chain)
            const double f = m.f;
            const double fp01 = fma(f, t.p11, t.p01);
            const double pp00 = fma(fp01, f, fma(f, t.p01, t.p00)) + m.q00;
            const double pp01 = fp01 + m.q01;
            const double pp11 = t.p11 + m.q11;
            const double s_ = pp00 + m.r_est;
            const double sinv = KFBO_RCP(s_);
            const double k0 = pp00 * sinv, k1 = pp01 * sinv;
            t.p00 = m.r_est * k0; t.p01 = m.r_est * k1; t.p11 = fma(-k1, pp01, pp11);
        }
        w = -w;
    }
    const long long c1 = clock64();
    double o[4];
    traj_finish(t, steps, o);
    if (o[0] + o[1] + t.p00 == 123.456) out[0] = o[0];
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[MODE] = c1 - c0;
}

int main()
{
    Case c{};
    c.f = 0.1; c.q00 = 3.3e-4; c.q01 = c.q10 = 5e-3; c.q11 = 0.1; c.r_est = 1.0; c.l00 = 0.018; c.l10 = 0.27; c.l11 = 0.16; c.sr = 1.0; c.T = 1000;
    double *d; long long *cyc;
    cudaMalloc(&d, 64); cudaMallocManaged(&cyc, 64);
    const int steps = 4000;
    for (int threads : {32, 128}) {
        k<0><<<1, threads>>>(c, d, steps, cyc); cudaDeviceSynchronize();
        k<0><<<1, threads>>>(c, d, steps, cyc); cudaDeviceSynchronize();
        k<1><<<1, threads>>>(c, d, steps, cyc); cudaDeviceSynchronize();
        std::printf("%3d threads in the CTA: full filter_step %.1f cycles per step; covariance recursion alone %.1f\n", threads, (double)cyc[0] / steps, (double)cyc[1] / steps);
    }
    return 0;
}
