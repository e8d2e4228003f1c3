// Micro-benchmark: throughput of float <-> double conversions (F2F on the XU pipe) next to DFMA, and of an
// integer-ALU float -> double widening, per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cvt_pipes cvt_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double widen_int(float f) {
    // exact for normal floats and zero: sign | (exp + 896) << 52 | mantissa << 29
    const unsigned u = __float_as_uint(f);
    const unsigned hi = (u & 0x80000000u) | (((u >> 3) & 0x0FFFFFFFu) + ((u & 0x7F800000u) ? 0x38000000u : 0u));
    return __hiloint2double((int)hi, (int)(u << 29));
}

template <int MODE, int CHAINS>
__global__ void cvt_kernel(double* out, int iters, double a, double b) {
    double v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = (double)(threadIdx.x + c) * 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (MODE == 0) v[c] = v[c] * a + b;                                   // DFMA
                if (MODE == 1) v[c] = (double)((float)v[c]) * a + b;                  // DFMA + F2F.F32.F64 + F2F.F64.F32
                if (MODE == 2) v[c] = widen_int((float)v[c]) * a + b;                 // DFMA + F2F.F32.F64 + integer widening
                if (MODE == 3) v[c] = widen_int(__int_as_float(__double2hiint(v[c]))) * a + b;   // DFMA + integer widening only
            }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int CHAINS>
void run(const char* name, int warps_per_sm, int sms) {
    double* out;
    cudaMalloc(&out, sizeof(double) * 1024 * 1024);
    const int iters = 10000;
    dim3 grid(sms), block(32 * warps_per_sm);
    cvt_kernel<MODE, CHAINS><<<grid, block>>>(out, 100, 1.0000001, 1e-9);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cvt_kernel<MODE, CHAINS><<<grid, block>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk_khz * 1e3;
    double steps = (double)iters * 8 * CHAINS * warps_per_sm;
    printf("%-44s warps/SM=%2d : %.2f SM-cycles per warp-step\n", name, warps_per_sm, cycles / steps);
    cudaFree(out);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {8, 16}) {
        if (w == 8) {
            run<0, 8>("DFMA", 8, sms);
            run<1, 8>("DFMA + F2F f64->f32 + F2F f32->f64", 8, sms);
            run<2, 8>("DFMA + F2F f64->f32 + integer widen", 8, sms);
            run<3, 8>("DFMA + integer widen", 8, sms);
        } else {
            run<0, 8>("DFMA", 16, sms);
            run<1, 8>("DFMA + F2F f64->f32 + F2F f32->f64", 16, sms);
            run<2, 8>("DFMA + F2F f64->f32 + integer widen", 16, sms);
            run<3, 8>("DFMA + integer widen", 16, sms);
        }
    }
    return 0;
}
