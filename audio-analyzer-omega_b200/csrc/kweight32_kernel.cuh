// kweight32_kernel.cuh -- the batch path's K-weighting + mean square in float32 state.
//
// Same contract and the same mapping as kweight_kernel.cuh (one warp per meter frame, 65 samples per lane
// in registers, zero-state sub-chunk sweeps + Kogge-Stone scan of the chunk states + homogeneous
// correction; ProfessionalMetering.apply_k_weighting, omega4/panels/professional_meters.py:129-153, and
// the mean square of calculate_lufs :237-246), for the case the batch entry point runs 11.5 M times per
// step: float32 samples times the Hann table, the K-weighting program of two Butterworth high-passes
// (b = b0 [1, -2, 1]) blended as f + 0.3 (s - f).
//
// float64 state (kweight_kernel.cuh) costs a 255-register kernel at 8 warps per SM that ncu shows bound by
// fixed-latency fp64 chains and ~200 F2F conversions per frame on the 16-lane XU pipe.  A direct-form
// recursion cannot simply be demoted to float32: the 38 Hz section has a double pole at radius 0.9965, its
// coefficients -1.99297 / 0.99299 lose the cut-off in float32, and round-off in the state is amplified by
// ~1 / (1 - r)^2.  The recursion is therefore restated so that every float32 quantity is either small or
// smooth (all constants derived in float64 on the host):
//
//   y[n] = b0 x[n] + z[n]                         z = the (low-frequency) part the high-pass removes
//   d[n] = a2 d[n-1] - alpha y[n-1] - b0 beta (x[n-1] - x[n-2])        alpha = 1 + a1 + a2 = A(1)
//   z[n] = z[n-1] + d[n]                                               beta  = 1 - a2
//
// (substitute y = b0 x + z into y[n] = b0 (x[n] - 2 x[n-1] + x[n-2]) - a1 y[n-1] - a2 y[n-2] and write the
// z recursion in delta form, d = z[n] - z[n-1]).  The kernel carries delta = d / (-b0 beta) instead of d,
// which turns the input term into a plain difference: 5 float32 operations per sample and sweep
//   delta = a2 delta' + (x' - x'') - gamma y',  z = z' + bb delta,  y = b0 x + z      (gamma = alpha / bb).  alpha = 2.47e-5 and beta = 7.0e-3 are stored to float32
// RELATIVE precision, so the pole positions keep 7 digits; for content above the cut-off z and d are
// ~beta |x|, for content below it z ~ -b0 x is smooth and d tiny, so the round-off that the double pole
// integrates is two orders of magnitude below that of a float32 direct form.  Measured against the float64
// oracle (numpy emulation of this exact structure: tests/tools/kweight32_numerics.py, profiles/r01h_kweight32_numerics.txt; on the GPU: tests/test_gpu_parity.py): <= 5e-6 LU on windowed
// frames of noise, tones 10 Hz .. 10 kHz, DC offsets, clipping, impulses and steps -- the float64 kernel's
// own level -- against the 0.01 LU bar.
//
// State of the scan is (z, delta); its one-step transition is M = [[1 - alpha, bb a2], [-gamma, a2]].  filtfilt's
// steady-state initial condition (lfilter_zi * x0) is, for a high-pass, "past inputs = x0, output 0":
// (z, d) = (-b0 x0, 0), and it is preserved over any run of constant input, so the 14 slack slots of lane
// 31 are simply filled with the last sample instead of moving the start back in time.
#pragma once
#include <cuda_runtime.h>
#include "kweight_kernel.cuh"

namespace o4 {

struct Kw32Sec {
    float b0, a2, gamma, bb;        // bb = -b0 beta, gamma = alpha / bb
    float phi[5][4];                // (M^65)^(2^j), row major
    float c16[4], c17[4];           // M^16, M^17
    float g[KW_SUBMAX][2];          // first row of M^(i+1)
};

struct Kweight32Args {
    const float* x;
    long long ch_stride, frame_stride, frame_off0;
    int n_ch, n_frames, first_frame, frames_per_warp;
    const float* hann;              // [W] float32 Hann
    double* lufs_out;               // [n_ch][n_frames]
    int rms_gate, _align;
    Kw32Sec s[2];
};

__device__ __forceinline__ float2 kw32_mat(const float* m, float2 s) {
    return make_float2(fmaf(m[0], s.x, m[1] * s.y), fmaf(m[2], s.x, m[3] * s.y));
}

template <bool BACKWARD>
__device__ __forceinline__ void kw32_pass(float (&r)[KW_L], const Kw32Sec& c, int lane) {
    const int pos = BACKWARD ? 31 - lane : lane;
    const float b0 = c.b0, a2 = c.a2, ngamma = -c.gamma, bb = c.bb;
    // the two inputs preceding each sub-chunk, as (x[-1], x[-1] - x[-2]); sub-chunk m in processing order
    float xa[KW_NSUB], dxa[KW_NSUB];
    {
        const float p1 = BACKWARD ? __shfl_down_sync(0xffffffffu, r[0], 1) : __shfl_up_sync(0xffffffffu, r[KW_L - 1], 1);
        const float p2 = BACKWARD ? __shfl_down_sync(0xffffffffu, r[1], 1) : __shfl_up_sync(0xffffffffu, r[KW_L - 2], 1);
        const float x0 = BACKWARD ? r[KW_L - 1] : r[0];                 // steady state: past inputs = first sample
        xa[0] = pos == 0 ? x0 : p1;
        dxa[0] = pos == 0 ? 0.f : p1 - p2;
    }
#pragma unroll
    for (int m = 1; m < KW_NSUB; ++m) {
        const int j = BACKWARD ? KW_NSUB - 1 - m : m;
        xa[m] = BACKWARD ? r[kw_off(j + 1)] : r[kw_off(j) - 1];
        dxa[m] = xa[m] - (BACKWARD ? r[kw_off(j + 1) + 1] : r[kw_off(j) - 2]);
    }
    const float2 s0 = make_float2(-b0 * xa[0], 0.f);                     // true entry state of the sequence (pos 0 only)
    // 1. zero-state sweeps, four independent chains
    float z1[KW_NSUB], d1[KW_NSUB], y1[KW_NSUB];
#pragma unroll
    for (int m = 0; m < KW_NSUB; ++m) { z1[m] = 0.f; d1[m] = 0.f; y1[m] = b0 * xa[m]; }
#pragma unroll
    for (int n = 0; n < KW_SUBMAX; ++n) {
#pragma unroll
        for (int m = 0; m < KW_NSUB; ++m) {
            const int j = BACKWARD ? KW_NSUB - 1 - m : m;
            const int len = kw_off(j + 1) - kw_off(j);
            if (n < len) {
                const int i = BACKWARD ? kw_off(j + 1) - 1 - n : kw_off(j) + n;
                const float x = r[i];
                const float t = fmaf(a2, d1[m], dxa[m]);
                const float d = fmaf(ngamma, y1[m], t);
                const float z = fmaf(bb, d, z1[m]);
                const float y = fmaf(b0, x, z);
                dxa[m] = x - xa[m]; xa[m] = x;
                z1[m] = z; d1[m] = d; y1[m] = y;
                r[i] = y;
            }
        }
    }
    // lane aggregate: zero-state end state of the whole 65-sample chunk
    float2 v = make_float2(z1[0], d1[0]);
#pragma unroll
    for (int m = 1; m < KW_NSUB; ++m) {
        const int j = BACKWARD ? KW_NSUB - 1 - m : m;
        const float* cm = (kw_off(j + 1) - kw_off(j) == 17) ? c.c17 : c.c16;
        const float2 q = kw32_mat(cm, v);
        v = make_float2(q.x + z1[m], q.y + d1[m]);
    }
    // 2. scan of the chunk end states across lanes
    if (pos == 0) {
        const float2 q = kw32_mat(c.phi[0], s0);
        v.x += q.x; v.y += q.y;
    }
#pragma unroll
    for (int jj = 0; jj < 5; ++jj) {
        const int dd = 1 << jj;
        const float rx = BACKWARD ? __shfl_down_sync(0xffffffffu, v.x, dd) : __shfl_up_sync(0xffffffffu, v.x, dd);
        const float ry = BACKWARD ? __shfl_down_sync(0xffffffffu, v.y, dd) : __shfl_up_sync(0xffffffffu, v.y, dd);
        if (pos >= dd) {
            const float2 q = kw32_mat(c.phi[jj], make_float2(rx, ry));
            v.x += q.x; v.y += q.y;
        }
    }
    float2 sin[KW_NSUB];
    sin[0].x = BACKWARD ? __shfl_down_sync(0xffffffffu, v.x, 1) : __shfl_up_sync(0xffffffffu, v.x, 1);
    sin[0].y = BACKWARD ? __shfl_down_sync(0xffffffffu, v.y, 1) : __shfl_up_sync(0xffffffffu, v.y, 1);
    if (pos == 0) sin[0] = s0;
#pragma unroll
    for (int m = 1; m < KW_NSUB; ++m) {
        const int jp = BACKWARD ? KW_NSUB - m : m - 1;
        const float* cm = (kw_off(jp + 1) - kw_off(jp) == 17) ? c.c17 : c.c16;
        const float2 q = kw32_mat(cm, sin[m - 1]);
        sin[m] = make_float2(q.x + z1[m - 1], q.y + d1[m - 1]);
    }
    // 3. homogeneous correction (y = b0 x + z: the state only enters through z)
#pragma unroll
    for (int n = 0; n < KW_SUBMAX; ++n) {
#pragma unroll
        for (int m = 0; m < KW_NSUB; ++m) {
            const int j = BACKWARD ? KW_NSUB - 1 - m : m;
            const int len = kw_off(j + 1) - kw_off(j);
            if (n < len) {
                const int i = BACKWARD ? kw_off(j + 1) - 1 - n : kw_off(j) + n;
                r[i] = fmaf(c.g[n][0], sin[m].x, fmaf(c.g[n][1], sin[m].y, r[i]));
            }
        }
    }
}

// odd reflection padding (9 samples) around the frame held at ext positions [9, 2057); the slack of lane 31
// continues the last padded sample (constant input keeps the steady state, see the header)
__device__ __forceinline__ void kw32_odd_pad(float (&r)[KW_L], int lane) {
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < KW_PAD; ++j) r[j] = 2.f * r[KW_PAD] - r[2 * KW_PAD - j];
    }
    if (lane == 31) {
        constexpr int E = KW_LAST - KW_PAD;
#pragma unroll
        for (int j = 0; j < KW_PAD; ++j) r[E + 1 + j] = 2.f * r[E] - r[E - 1 - j];
#pragma unroll
        for (int i = KW_LAST + 1; i < KW_L; ++i) r[i] = r[KW_LAST];
    }
}

__device__ __forceinline__ void kw32_fill_slack(float (&r)[KW_L], int lane) {
    if (lane == 31) {
#pragma unroll
        for (int i = KW_LAST + 1; i < KW_L; ++i) r[i] = r[KW_LAST];
    }
}

constexpr int KW32_WARPS = 4;
#ifndef KW32_MIN_BLOCKS
#define KW32_MIN_BLOCKS 3
#endif

__global__ void __launch_bounds__(KW32_WARPS * 32, KW32_MIN_BLOCKS)
kweight32_kernel(const __grid_constant__ Kweight32Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* hann_x = reinterpret_cast<float*>(smem_raw);                 // [32*65] ext layout, 0 at the pads
    float* stage_all = hann_x + 32 * KW_L;                              // [KW32_WARPS][KW_STG]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* stg = stage_all + warp * KW_STG;

    for (int p = threadIdx.x; p < 32 * KW_L; p += blockDim.x) {
        const int n = p - KW_PAD;
        hann_x[p] = (n >= 0 && n < KW_W) ? a.hann[n] : 0.f;
    }
    __syncthreads();

    const int frames_per_cta = a.frames_per_warp * KW32_WARPS;
    const int tiles_per_ch = (a.n_frames + frames_per_cta - 1) / frames_per_cta;
    const int ch = blockIdx.x / tiles_per_ch;
    const int tile = blockIdx.x % tiles_per_ch;

    for (int it = 0; it < a.frames_per_warp; ++it) {
        const int f = tile * frames_per_cta + it * KW32_WARPS + warp;
        if (f >= a.n_frames || f < a.first_frame) continue;             // warp-uniform
        const long long off = (long long)ch * a.ch_stride + a.frame_off0 + (long long)f * a.frame_stride;

        // coalesced copy of the frame into the staging strip (pads zeroed), then conflict-free strided reads
        const float4* px = reinterpret_cast<const float4*>(a.x + off);
        float4* s4 = reinterpret_cast<float4*>(stg + KW_PAD + KW_STG_SHIFT);
        float4 tmp[KW_W / 4 / 32];
#pragma unroll
        for (int j = 0; j < KW_W / 4 / 32; ++j) tmp[j] = __ldg(px + lane + 32 * j);
        if (lane < KW_PAD) stg[KW_STG_SHIFT + lane] = 0.f;
        if (lane < KW_SLACK + KW_PAD + KW_STG_SHIFT) stg[KW_STG_SHIFT + KW_PAD + KW_W + lane] = 0.f;
#pragma unroll
        for (int j = 0; j < KW_W / 4 / 32; ++j) s4[lane + 32 * j] = tmp[j];
        __syncwarp();
        float* sl = stg + KW_STG_SHIFT + KW_L * lane;
        const float* hl = hann_x + KW_L * lane;
        float r[KW_L];
        float sumsq = 0.f;
#pragma unroll
        for (int i = 0; i < KW_L; ++i) {
            const float xv = sl[i] * hl[i];
            sumsq = fmaf(xv, xv, sumsq);
            r[i] = xv;
        }
        __syncwarp();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
        const bool gated = a.rms_gate && sqrtf(sumsq * (1.f / KW_W)) < 1e-6f;   // professional_meters.py:132-134

        double ms = 0.0;
        if (!gated) {                                                    // warp-uniform
            // section loop not unrolled: one forward and one backward body serve both sections
#ifdef KW32_UNROLL_SECTIONS
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int si = 0; si < 2; ++si) {
                const Kw32Sec& c = a.s[si];
                kw32_odd_pad(r, lane);
                kw32_pass<false>(r, c, lane);
                kw32_fill_slack(r, lane);
                kw32_pass<true>(r, c, lane);
                if (si == 0) {
                    // zero the pad positions and stash f (first filtfilt output) in the staging strip
                    kw_zero_pads(r, lane);
#pragma unroll
                    for (int i = 0; i < KW_L; ++i) sl[i] = r[i];
                }
            }
            kw_zero_pads(r, lane);
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < KW_L; ++i) {
                const float fv = sl[i];
                const float w = fmaf(r[i] - fv, 0.3f, fv);               // f + (s - f) * 0.3 ; 0 at the pads
                acc = fmaf(w, w, acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            ms = (double)acc / (double)KW_W;
            __syncwarp();
        }
        if (lane == 0)
            a.lufs_out[(size_t)ch * a.n_frames + f] = (ms > 1e-10) ? (-0.691 + 10.0 * log10(ms)) : -100.0;
    }
}

inline size_t kweight32_smem_bytes() {
    return (size_t)32 * KW_L * sizeof(float) + (size_t)KW32_WARPS * KW_STG * sizeof(float);
}

}  // namespace o4
