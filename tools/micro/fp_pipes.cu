// Micro-benchmark: fp64 / fp32 FMA latency and throughput per SM on this GPU.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp_pipes fp_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int CHAINS>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
    T v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = (T)(threadIdx.x + c);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) v[c] = v[c] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename T, int CHAINS>
void run(const char* name, int warps_per_sm, int sms) {
    T* out;
    cudaMalloc(&out, sizeof(T) * 1024 * 1024 * 4);
    const int iters = 20000;
    dim3 grid(sms), block(32 * warps_per_sm);
    fma_kernel<T, CHAINS><<<grid, block>>>(out, 100, (T)1.0000001, (T)1e-9);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    fma_kernel<T, CHAINS><<<grid, block>>>(out, iters, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk_khz * 1e3;
    double fma_per_warp = (double)iters * 8 * CHAINS;
    printf("%-6s chains=%2d warps/SM=%2d : %.2f cycles per FMA per warp, %.2f warp-FMA/cycle/SM (%.1f lanes/clk/SM) [clk %d MHz nominal]\n",
           name, CHAINS, warps_per_sm, cycles / fma_per_warp, fma_per_warp * warps_per_sm / cycles,
           32.0 * fma_per_warp * warps_per_sm / cycles, clk_khz / 1000);
    cudaFree(out);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    run<double, 1>("fp64", 4, sms);     // 1 warp per scheduler, dependent chain -> latency
    run<double, 2>("fp64", 4, sms);
    run<double, 4>("fp64", 4, sms);
    run<double, 8>("fp64", 4, sms);
    run<double, 4>("fp64", 8, sms);
    run<double, 8>("fp64", 8, sms);
    run<double, 8>("fp64", 16, sms);
    run<double, 8>("fp64", 32, sms);
    run<float, 1>("fp32", 4, sms);
    run<float, 4>("fp32", 4, sms);
    run<float, 8>("fp32", 4, sms);
    run<float, 8>("fp32", 16, sms);
    run<float, 8>("fp32", 32, sms);
    return 0;
}
