// rng_only.cu -- SMSP-cycles per warp for the pieces of the noise generator:
//   (a) Philox4x32-10 alone  (b) normal_pair alone (bits from a cheap LCG)  (c) both
//   (d) IMAD.WIDE.U32 and MUFU.RSQ64H / RCP64H issue rates
#include <cstdio>
#include <cuda_runtime.h>
#include "../../kf_bayesopt_b200/csrc/kfbo_device.cuh"

using namespace kfbo;

template <int MODE>
__global__ void __launch_bounds__(128) k(const math::LogTables *logtab, const PhiloxKeys key, double *out, int iters)
{
    extern __shared__ __align__(16) unsigned char s_raw[];     // the noise tables (48 KB): dynamic shared memory
    math::LogTables &s_log = *reinterpret_cast<math::LogTables *>(s_raw);
    {
        const double *ls = reinterpret_cast<const double *>(logtab);
        double *ld = reinterpret_cast<double *>(&s_log);
        for (int i = threadIdx.x; i < (int)(sizeof(math::LogTables) / sizeof(double)); i += 128) ld[i] = ls[i];
    }
    __syncthreads();
    const uint32_t trial = blockIdx.x * 128 + threadIdx.x;
    double acc = 0;
    uint32_t iacc = 0;
    uint4 r = make_uint4(trial * 2654435761u, trial ^ 0x9E3779B9u, trial + 0x85EBCA6Bu, ~trial);
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {   // 3 Philox calls
#pragma unroll
            for (int c = 0; c < 3; ++c) { uint4 q = philox4x32_10(trial, 7u, (uint32_t)i, (uint32_t)c, key); iacc ^= q.x ^ q.y ^ q.z ^ q.w; }
        } else if (MODE == 1) {   // 3 normal pairs from cheap bits
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                r.x = r.x * 1664525u + 1013904223u; r.y ^= r.x >> 3; r.z += r.y; r.w ^= r.z << 5;
                double a, b;
                math::normal_pair(r.x ^ r.z, r.y ^ r.w, s_log, a, b);
                acc += a * b;
            }
        } else if (MODE == 2) {   // both
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double a, b;
                const uint4 q = philox4x32_10(trial, 7u, (uint32_t)i, (uint32_t)c, key);
                math::normal_pair(q.x ^ q.z, q.y ^ q.w, s_log, a, b);
                acc += a * b;
            }
        } else if (MODE == 3) {   // 24 IMAD.WIDE.U32 (8 independent chains x 3)
            uint32_t y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = r.x + j;
#pragma unroll
            for (int u = 0; u < 3; ++u)
#pragma unroll
                for (int j = 0; j < 8; ++j) { const uint64_t p = (uint64_t)y[j] * 0xD2511F53u; y[j] = (uint32_t)p ^ (uint32_t)(p >> 32); }
#pragma unroll
            for (int j = 0; j < 8; ++j) iacc ^= y[j];
            r.x = iacc;
        } else if (MODE == 4) {   // 8 MUFU.RSQ64H + 8 MUFU.RCP64H, independent
            double z[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) z[j] = acc + 1.5 + j;
#pragma unroll
            for (int j = 0; j < 8; ++j) z[j] = math::rsqrt_seed(z[j]) + math::rcp_seed(z[j]);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += z[j];
        }
    }
    if (acc == 123.456 || iacc == 0xdeadbeef) out[0] = acc + iacc;
}

template <int MODE>
void run(const math::LogTables *dt, double *d, const char *name, int blocks_per_sm)
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * blocks_per_sm, iters = 3000;
    const PhiloxKeys key = {{1, 2, 3, 4, 5, 6, 7, 8, 9, 10}, {11, 12, 13, 14, 15, 16, 17, 18, 19, 20}};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(math::LogTables));
    k<MODE><<<blocks, 128, sizeof(math::LogTables)>>>(dt, key, d, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 128, sizeof(math::LogTables)>>>(dt, key, d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_iters_per_smsp = (double)iters * (blocks * 4) / (p.multiProcessorCount * 4);
    std::printf("%-44s blocks/SM=%d : %.1f SMSP-cycles per warp-iteration\n", name, blocks_per_sm, ms * 1e-3 * p.clockRate * 1e3 / warp_iters_per_smsp);
}

int main()
{
    static math::LogTables t;
    math::build_log_tables(&t);
    math::LogTables *dt; double *d;
    cudaMalloc(&dt, sizeof t); cudaMalloc(&d, 64);
    cudaMemcpy(dt, &t, sizeof t, cudaMemcpyHostToDevice);
    for (int b : {3, 5, 8}) {
        run<0>(dt, d, "3 x Philox4x32-10", b);
        run<1>(dt, d, "3 x normal_pair (cheap bits)", b);
        run<2>(dt, d, "3 x (Philox + normal_pair)", b);
        run<3>(dt, d, "24 x IMAD.WIDE.U32 (+24 LOP3)", b);
        run<4>(dt, d, "8 x MUFU.RSQ64H + 8 x MUFU.RCP64H (+16 DADD)", b);
    }
    return 0;
}
