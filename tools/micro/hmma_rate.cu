// Micro-benchmark: issue rate of the warp-level mma.sync.m16n8k16 (f16 inputs; f32 or f16 accumulators) on sm_100a --
// the legacy tensor path (HMMA), usable from ordinary warps without TMEM.  Dense MACs per clock and SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int ACC32, int NACC>
__global__ void k(float* out, int iters) {
    uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x38003800u, 0x34003400u};
    uint32_t b[2] = {0x3c003c00u, 0x2c002c00u + threadIdx.x};
    float c[NACC][4];
    uint32_t h[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; h[i][0] = h[i][1] = 0u; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (ACC32)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
                             : "+r"(h[i][0]), "+r"(h[i][1])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc += c[i][0] + c[i][1] + c[i][2] + c[i][3] + (float)h[i][0] + (float)h[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int ACC32, int NACC>
void run(const char* name, int warps, int sms) {
    float* out; cudaMalloc(&out, 4 << 22);
    const int iters = 20000;
    k<ACC32, NACC><<<sms, 32 * warps>>>(out, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ACC32, NACC><<<sms, 32 * warps>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * khz * 1e3;
    const double mmas = (double)iters * NACC * warps;               // per SM
    printf("%-40s warps/SM %2d, %d independent accumulators: %.3f MMA/clk/SM = %.0f MACs/clk/SM (%.1f clk per MMA and scheduler)\n",
           name, warps, NACC, mmas / cyc, mmas / cyc * 2048, cyc / (mmas / 4));
    cudaFree(out);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<1, 4>("mma.sync m16n8k16 f16 -> f32", 16, sms);
    run<1, 8>("mma.sync m16n8k16 f16 -> f32", 16, sms);
    run<1, 8>("mma.sync m16n8k16 f16 -> f32", 8, sms);
    run<1, 8>("mma.sync m16n8k16 f16 -> f32", 4, sms);
    run<0, 8>("mma.sync m16n8k16 f16 -> f16", 16, sms);
    run<0, 8>("mma.sync m16n8k16 f16 -> f16", 4, sms);
    return 0;
}
