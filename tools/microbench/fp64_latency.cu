// fp64_latency.cu -- dependent-issue latency of DFMA/DADD/DMUL/MUFU.RCP64H and FP64 throughput vs
// resident warps x independent chains per warp (how much parallelism saturates the FP64 pipe).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat(double *out, long long *cyc, double a, double b)
{
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 4096; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    double y = x;
#pragma unroll 64
    for (int i = 0; i < 4096; ++i) y = y + a;
    long long t2 = clock64();
    double z = y;
#pragma unroll 64
    for (int i = 0; i < 4096; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z)); z = r; }
    long long t3 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
    if (z == 123.456) out[0] = z;
}

template <int CHAINS>
__global__ void thr(double *out, int iters, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) x[j] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) x[j] = fma(x[j], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) s += x[j];
    if (s == 123.456) out[0] = s;
}

template <int CHAINS>
void run_thr(double *d, int sms, int warps_per_smsp)
{
    const int threads = warps_per_smsp * 4 * 32;   // one CTA per SM
    const int iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    thr<CHAINS><<<sms, threads>>>(d, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    thr<CHAINS><<<sms, threads>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double instr_per_smsp = (double)iters * 16 * CHAINS * warps_per_smsp;
    const double cyc = ms * 1e-3 * p.clockRate * 1e3;
    std::printf("warps/SMSP=%d chains=%d : %.2f cycles per DFMA per SMSP (2.0 = peak)\n", warps_per_smsp, CHAINS, cyc / instr_per_smsp);
}

int main()
{
    double *d; long long *c, h[3];
    cudaMalloc(&d, 64); cudaMalloc(&c, 64);
    lat<<<1, 32>>>(d, c, 0.999999, 1e-9);
    cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
    std::printf("dependent latency (cycles): DFMA %.2f  DADD %.2f  MUFU.RCP64H %.2f\n", h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    for (int w : {1, 2, 3, 4, 5, 8}) { run_thr<1>(d, sms, w); run_thr<2>(d, sms, w); run_thr<4>(d, sms, w); }
    return 0;
}
