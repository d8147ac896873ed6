// filter_only.cu -- cycles per filter_step (no RNG, no memory traffic) for K trajectories per thread
// and the Case constants in vector registers (loaded from global) vs uniform (kernel parameter).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../kf_bayesopt_b200/csrc/kfbo_device.cuh"

using namespace kfbo;

template <int K, bool PARAM>
__global__ void __launch_bounds__(128) k(const Case *cases, const Case pc, double *out, int steps)
{
    const Case m = PARAM ? pc : cases[blockIdx.x & 1];
    Traj t[K];
#pragma unroll
    for (int j = 0; j < K; ++j) traj_init(t[j], 0.01 * threadIdx.x, 0.02 * j);
    double w = 1e-3 * threadIdx.x;
    for (int s = 0; s < steps; ++s) {
        const double2 g = make_double2(1e-3 * (s & 7), 1e-2 * (s & 3));
#pragma unroll
        for (int j = 0; j < K; ++j) filter_step<false, false>(t[j], m, m.f, g, StepNoise{w, -w, 0.5 * w});
        w = -w;
    }
    double o[4], acc = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) { traj_finish(t[j], steps, o); acc += o[0] + o[1] + o[2] + o[3]; }
    if (acc == 123.456) out[0] = acc;
}

template <int K, bool PARAM>
void run(const Case *d_cases, const Case &c, double *d, int blocks_per_sm)
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * blocks_per_sm, steps = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<K, PARAM><<<blocks, 128>>>(d_cases, c, d, steps);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<K, PARAM><<<blocks, 128>>>(d_cases, c, d, steps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_steps_per_smsp = (double)steps * K * (blocks * 4) / (p.multiProcessorCount * 4);
    std::printf("K=%d %s blocks/SM=%d : %.1f SMSP-cycles per warp-step, %.3e steps/s\n", K, PARAM ? "param " : "global", blocks_per_sm,
                ms * 1e-3 * p.clockRate * 1e3 / warp_steps_per_smsp, (double)steps * K * blocks * 128 / (ms * 1e-3));
}

int main()
{
    Case c{};
    c.f = 0.1; c.q00 = 3.3e-4; c.q01 = c.q10 = 5e-3; c.q11 = 0.1; c.r_est = 1.0; c.l00 = 0.018; c.l10 = 0.27; c.l11 = 0.16; c.sr = 1.0; c.T = 1000;
    Case *dc; double *d;
    cudaMalloc(&dc, 2 * sizeof(Case)); cudaMalloc(&d, 64);
    Case two[2] = {c, c};
    cudaMemcpy(dc, two, sizeof two, cudaMemcpyHostToDevice);
    for (int b : {2, 4, 8}) {
        run<1, false>(dc, c, d, b); run<1, true>(dc, c, d, b);
        run<2, false>(dc, c, d, b); run<2, true>(dc, c, d, b);
        run<4, true>(dc, c, d, b);
    }
    return 0;
}
