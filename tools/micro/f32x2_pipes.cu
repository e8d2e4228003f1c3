// Micro-benchmark: packed fp32 (add.f32x2 / fma.rn.f32x2, sm_100) against scalar FADD / FFMA -- pipe rate, and
// whether the saved issue slots are usable by other instructions (here: shared-memory loads) next to them.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_pipes f32x2_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int MODE, int LDS_PER_8>
__global__ void k(float* out, int iters) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (float)i * 1e-9f;
    __syncthreads();
    float2 v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = make_float2(threadIdx.x + c, threadIdx.x - c);
    const float2 a = make_float2(1e-9f, 2e-9f), m = make_float2(1.0000001f, 0.9999999f);
    float acc = 0.f;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            if (MODE == 0) { v[c].x += a.x; v[c].y += a.y; }                       // 2 FADD
            if (MODE == 1) v[c] = __fadd2_rn(v[c], a);                              // 1 FADD2
            if (MODE == 2) { v[c].x = fmaf(v[c].x, m.x, a.x); v[c].y = fmaf(v[c].y, m.y, a.y); }   // 2 FFMA
            if (MODE == 3) v[c] = __ffma2_rn(v[c], m, a);                           // 1 FFMA2
            if (LDS_PER_8 > 0 && (c % (8 / LDS_PER_8)) == 0 && c < 8) { acc += sm[idx & 4095]; idx += 33; }
        }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) acc += v[c].x + v[c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// half2: one HADD2 / HFMA2 per element and iteration (two half lanes each)
template <int MODE>
__global__ void kh(float* out, int iters) {
    __half2 v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = __floats2half2_rn(threadIdx.x * 1e-3f + c, threadIdx.x * 1e-3f - c);
    const __half2 a = __floats2half2_rn(1e-3f, 2e-3f), m = __floats2half2_rn(1.0009f, 0.9991f);
    float f[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) f[c] = threadIdx.x * 1e-3f + c;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 6) {            // 16 HFMA2 + 16 FFMA interleaved: do float32 and half2 share the pipe?
#pragma unroll
            for (int c = 0; c < 16; ++c) { v[c] = __hfma2(v[c], m, a); f[c] = fmaf(f[c], 1.0000001f, 1e-9f); }
        }
        if (MODE == 7) {            // 16 HADD2 + 16 FFMA interleaved
#pragma unroll
            for (int c = 0; c < 16; ++c) { v[c] = __hadd2(v[c], a); f[c] = fmaf(f[c], 1.0000001f, 1e-9f); }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            if (MODE == 0) v[c] = __hadd2(v[c], a);
            if (MODE == 1) v[c] = __hfma2(v[c], m, a);
            if (MODE == 2) v[c] = __hmul2(v[c], m);
            if (MODE == 3) v[c] = (c & 1) ? __hadd2(v[c], a) : __hfma2(v[c], m, a);          // HADD2 / HFMA2 alternating
            if (MODE == 4) v[c] = (c & 1) ? __hmul2(v[c], m) : __hfma2(v[c], m, a);          // HMUL2 / HFMA2 alternating
            if (MODE == 5) v[c] = (c & 1) ? __hadd2(v[c], a) : __hmul2(v[c], m);             // HADD2 / HMUL2 alternating
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) acc += __low2float(v[c]) + __high2float(v[c]);
#pragma unroll
    for (int c = 0; c < 16; ++c) acc += f[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE>
void runh(const char* name, int warps, int sms) {
    float* out; cudaMalloc(&out, 4 << 22);
    const int iters = 4000;
    kh<MODE><<<sms, 32 * warps>>>(out, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kh<MODE><<<sms, 32 * warps>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * khz * 1e3;
    const double inst_lane = (double)iters * 32 * warps * 32;     // 32 half2 instructions per thread and iteration
    printf("%-22s warps/SM %2d: %.1f half2 instruction lanes/clk/SM (= %.2f warp instructions/clk/SM)\n", name, warps, inst_lane / cyc, inst_lane / cyc / 32);
    cudaFree(out);
}

template <int MODE, int L>
void run(const char* name, int warps, int sms) {
    float* out; cudaMalloc(&out, 4 << 22);
    const int iters = 4000;
    k<MODE, L><<<sms, 32 * warps>>>(out, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, L><<<sms, 32 * warps>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * khz * 1e3;
    const double flops_lane = (double)iters * 32 * warps * 32;     // 32 scalar fp ops per thread and iteration
    printf("%-22s warps/SM %2d, %d LDS per 16 fp ops: %.1f fp lane-ops/clk/SM, %.0f cycles per iteration per scheduler-warp\n",
           name, warps, L, flops_lane / cyc, cyc / iters / (warps / 4.0));
    cudaFree(out);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0, 0>("FADD scalar", 16, sms);  run<1, 0>("FADD2 packed", 16, sms);
    run<2, 0>("FFMA scalar", 16, sms);  run<3, 0>("FFMA2 packed", 16, sms);
    run<0, 4>("FADD scalar + LDS", 16, sms);  run<1, 4>("FADD2 packed + LDS", 16, sms);
    run<2, 4>("FFMA scalar + LDS", 16, sms);  run<3, 4>("FFMA2 packed + LDS", 16, sms);
    run<0, 8>("FADD scalar + LDS", 16, sms);  run<1, 8>("FADD2 packed + LDS", 16, sms);
    run<0, 0>("FADD scalar", 8, sms);   run<1, 0>("FADD2 packed", 8, sms);
    runh<0>("HADD2", 16, sms); runh<1>("HFMA2", 16, sms); runh<2>("HMUL2", 16, sms); runh<1>("HFMA2", 8, sms);
    runh<3>("HADD2 / HFMA2 mix", 16, sms); runh<4>("HMUL2 / HFMA2 mix", 16, sms); runh<5>("HADD2 / HMUL2 mix", 16, sms);
    runh<6>("16 HFMA2 + 16 FFMA", 16, sms); runh<7>("16 HADD2 + 16 FFMA", 16, sms);
    return 0;
}
