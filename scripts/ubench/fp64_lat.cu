// FP64 pipe microbenchmark for B200: DFMA throughput per SM as a function of resident warps and
// independent chains per thread (ILP).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b) {
    double r[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) r[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) r[i] = fma(r[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += r[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int warps_per_sm, int sms, double *d) {
    const int iters = 2048;
    int threads = warps_per_sm <= 32 ? warps_per_sm * 32 : 1024, blocks = sms * (warps_per_sm <= 32 ? 1 : warps_per_sm / 32);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<blocks, threads>>>(d, 16, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dfma = (double)blocks * threads * iters * 16.0 * ILP;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("warps/SM %2d ILP %d : %.2f TFLOP/s, %.1f DFMA lanes/clk/SM (at %d MHz nominal)\n", warps_per_sm, ILP, 2 * dfma / ms / 1e9,
           dfma / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *d; cudaMalloc(&d, 1 << 24);
    for (int w : {4, 8, 12, 16, 32, 64}) { run<1>(w, sms, d); run<2>(w, sms, d); run<4>(w, sms, d); run<8>(w, sms, d); }
    return 0;
}
