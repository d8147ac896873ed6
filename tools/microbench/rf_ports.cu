// rf_ports.cu -- does a DFMA with three DISTINCT vector-register sources cost more than one with two?
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b)
{
    double x[8], y[8], z[8], w[8];
    const double c = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) { x[j] = threadIdx.x + j; y[j] = 1.0 + 1e-9 * (threadIdx.x + j); z[j] = 1e-9 * j * threadIdx.x; w[j] = 0.5 * threadIdx.x - j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (MODE == 0) x[j] = fma(x[j], a, b);            // 1 vector-register source
                if (MODE == 1) x[j] = fma(x[j], y[j], b);         // 2
                if (MODE == 2) x[j] = fma(x[j], y[j], z[j]);      // 3 distinct
                if (MODE == 3) x[j] = fma(x[j], y[j], x[j]);      // 3 slots, 2 distinct
                if (MODE == 4) x[j] = x[j] * y[j];                // DMUL 2
                if (MODE == 5) x[j] = x[j] + y[j];                // DADD 2
                if (MODE == 6) x[j] = fma(c, x[j], z[j]);         // 3 distinct, one of them the same register in every instruction (.reuse)
                if (MODE == 7) { x[j] = fma(x[j], y[j], z[j]); w[j] = fma(w[j], a, b); }   // 3 distinct interleaved with 1 source: do the register reads average out?
                if (MODE == 8) { x[j] = fma(x[j], y[j], z[j]); w[j] = fma(w[j], y[j], b); } // 3 distinct + 2 sources, the second re-reading y[j] in the same slot
            }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j] + y[j] + z[j] + w[j];
    if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(double *d, const char *name)
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 4, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double instr_per_smsp = (double)iters * (MODE >= 7 ? 128 : 64) * (blocks * 256 / 32) / (p.multiProcessorCount * 4);
    std::printf("%-40s %.2f cycles per instruction per SMSP\n", name, ms * 1e-3 * p.clockRate * 1e3 / instr_per_smsp);
}

int main()
{
    double *d;
    cudaMalloc(&d, 64);
    run<0>(d, "DFMA x = fma(x, U, U)   1 reg source");
    run<1>(d, "DFMA x = fma(x, y, U)   2 reg sources");
    run<2>(d, "DFMA x = fma(x, y, z)   3 distinct");
    run<3>(d, "DFMA x = fma(x, y, x)   2 distinct");
    run<4>(d, "DMUL x = x * y");
    run<5>(d, "DADD x = x + y");
    run<6>(d, "DFMA x = fma(c, x, z)   3 distinct, c reused");
    run<7>(d, "DFMA 3 distinct + DFMA 1 source, alternating");
    run<8>(d, "DFMA 3 distinct + DFMA 2 sources (y re-read)");
    return 0;
}
