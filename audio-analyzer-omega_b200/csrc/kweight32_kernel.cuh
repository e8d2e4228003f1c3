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
#include "blockdft_tc_kernel.cuh"      // hop_row_scale

namespace o4 {

struct Kw32Sec {
    float b0, a2, gamma, bb;        // bb = -b0 beta, gamma = alpha / bb
    float phi[5][4];                // (M^65)^(2^j), row major
    float c16[4], c17[4];           // M^16, M^17
    float g[KW_SUBMAX][2];          // first row of M^(i+1)
    float2 ga0[16], ga1[16];        // (g[k+1][e], g[k][e]), e = 0, 1: the forward correction of the (sub-chunk 0, sub-chunk 3) pair
};

struct Kweight32Args {
    const float* x;
    long long ch_stride, frame_stride, frame_off0;
    int n_ch, n_frames, first_frame, frames_per_warp;
    const float* hann;              // [W] float32 Hann
    double* lufs_out;               // [n_ch][n_frames]
    int rms_gate, _align;
    // optional by-product for the tensor-core hop-block GEMM of the same call (blockdft_tc_kernel.cuh): the operand scale
    // of the hop block each frame ends with (hop 512: the last 128 float4 of the frame, already in registers here);
    // hop_inv[ch][f - hop_inv_j0], row length hop_inv_nb.  nullptr: not wanted
    float* hop_inv;
    int hop_inv_nb, hop_inv_j0;
    Kw32Sec s[2];
};

__device__ __forceinline__ float2 kw32_mat(const float* m, float2 s) {
    return make_float2(fmaf(m[0], s.x, m[1] * s.y), fmaf(m[2], s.x, m[3] * s.y));
}

// The lane's 65 samples as packed-float32 operands (add.f32x2 / fma.rn.f32x2, sm_100): the four independent sub-chunk
// chains of a sweep run as TWO packed chains, so a step costs 5 issue slots per pair instead of 10 (the FMA pipe's lane
// rate is the same, tools/micro/f32x2_pipes.cu; the kernel was short of issue slots: issue 67 %, FMA pipe 50 %).
//   sub-chunk 0 = local samples 0 .. 16 (17), 1 = 17 .. 32, 2 = 33 .. 48, 3 = 49 .. 64 (16 each)
//   a[k] = (sample 1 + k, sample 49 + k)      sub-chunk 0 without its first sample, paired with sub-chunk 3
//   b[k] = (sample 17 + k, sample 33 + k)     sub-chunks 1 and 2
//   r0   = sample 0: one scalar step before (forward) or after (backward) the 16 packed steps of its chain
// Every element sees exactly the operations of the scalar formulation, in the same order: results are bit-identical
// (tests/tools/kweight32_numerics.py emulates that order).
struct Kw32R { float r0; float2 a[16]; float2 b[16]; };

__device__ __forceinline__ float& kwr(Kw32R& q, int i) {               // local sample i, i a compile-time constant after unrolling
    if (i == 0) return q.r0;
    if (i <= 16) return q.a[i - 1].x;
    if (i <= 32) return q.b[i - 17].x;
    if (i <= 48) return q.b[i - 33].y;
    return q.a[i - 49].y;
}
__device__ __forceinline__ float2 bc2(float v) { return make_float2(v, v); }

template <bool BACKWARD>
__device__ __forceinline__ void kw32_pass(Kw32R& q, const Kw32Sec& c, int lane) {
    const int pos = BACKWARD ? 31 - lane : lane;
    const float b0 = c.b0, a2 = c.a2, ngamma = -c.gamma, bb = c.bb;
    // the two inputs preceding each sub-chunk, as (x[-1], x[-1] - x[-2]), captured before anything is overwritten
    float xa_in, dxa_in;                                                // the chain that starts at the lane boundary
    {
        const float p1 = BACKWARD ? __shfl_down_sync(0xffffffffu, kwr(q, 0), 1) : __shfl_up_sync(0xffffffffu, kwr(q, KW_L - 1), 1);
        const float p2 = BACKWARD ? __shfl_down_sync(0xffffffffu, kwr(q, 1), 1) : __shfl_up_sync(0xffffffffu, kwr(q, KW_L - 2), 1);
        const float x0 = BACKWARD ? kwr(q, KW_L - 1) : kwr(q, 0);       // steady state: past inputs = first sample
        xa_in = pos == 0 ? x0 : p1;
        dxa_in = pos == 0 ? 0.f : p1 - p2;
    }
    const float2 s0 = make_float2(-b0 * xa_in, 0.f);                    // true entry state of the sequence (pos 0 only)
    // pair A = (sub-chunk 0, sub-chunk 3), pair B = (sub-chunk 1, sub-chunk 2)
    float2 xaA, dxaA, xaB, dxaB;
    if (!BACKWARD) {
        xaA = make_float2(xa_in, kwr(q, 48)); dxaA = make_float2(dxa_in, kwr(q, 48) - kwr(q, 47));
        xaB = make_float2(kwr(q, 16), kwr(q, 32)); dxaB = make_float2(kwr(q, 16) - kwr(q, 15), kwr(q, 32) - kwr(q, 31));
    } else {
        xaA = make_float2(kwr(q, 17), xa_in); dxaA = make_float2(kwr(q, 17) - kwr(q, 18), dxa_in);
        xaB = make_float2(kwr(q, 33), kwr(q, 49)); dxaB = make_float2(kwr(q, 33) - kwr(q, 34), kwr(q, 49) - kwr(q, 50));
    }
    // 1. zero-state sweeps
    float2 zA = bc2(0.f), dA = bc2(0.f), yA = make_float2(b0 * xaA.x, b0 * xaA.y);
    float2 zB = bc2(0.f), dB = bc2(0.f), yB = make_float2(b0 * xaB.x, b0 * xaB.y);
    if (!BACKWARD) {                                                    // sub-chunk 0's first sample
        const float x = q.r0;
        const float t = fmaf(a2, dA.x, dxaA.x);
        const float d = fmaf(ngamma, yA.x, t);
        const float z = fmaf(bb, d, zA.x);
        const float y = fmaf(b0, x, z);
        dxaA.x = x - xaA.x; xaA.x = x; zA.x = z; dA.x = d; yA.x = y;
        q.r0 = y;
    }
    const float2 A2 = bc2(a2), NG = bc2(ngamma), BB = bc2(bb), B0 = bc2(b0);
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        const int k = BACKWARD ? 15 - n : n;
        {
            const float2 x = q.a[k];
            const float2 t = __ffma2_rn(A2, dA, dxaA);
            const float2 d = __ffma2_rn(NG, yA, t);
            const float2 z = __ffma2_rn(BB, d, zA);
            const float2 y = __ffma2_rn(B0, x, z);
            dxaA = __fadd2_rn(x, make_float2(-xaA.x, -xaA.y)); xaA = x; zA = z; dA = d; yA = y;
            q.a[k] = y;
        }
        {
            const float2 x = q.b[k];
            const float2 t = __ffma2_rn(A2, dB, dxaB);
            const float2 d = __ffma2_rn(NG, yB, t);
            const float2 z = __ffma2_rn(BB, d, zB);
            const float2 y = __ffma2_rn(B0, x, z);
            dxaB = __fadd2_rn(x, make_float2(-xaB.x, -xaB.y)); xaB = x; zB = z; dB = d; yB = y;
            q.b[k] = y;
        }
    }
    if (BACKWARD) {                                                     // sub-chunk 0's last processed sample
        const float x = q.r0;
        const float t = fmaf(a2, dA.x, dxaA.x);
        const float d = fmaf(ngamma, yA.x, t);
        const float z = fmaf(bb, d, zA.x);
        const float y = fmaf(b0, x, z);
        zA.x = z; dA.x = d;
        q.r0 = y;
    }
    // zero-state end states per sub-chunk, in processing order m = 0 .. 3
    //   forward: sub-chunks 0, 1, 2, 3     backward: 3, 2, 1, 0
    float2 e[KW_NSUB];
    if (!BACKWARD) { e[0] = make_float2(zA.x, dA.x); e[1] = make_float2(zB.x, dB.x); e[2] = make_float2(zB.y, dB.y); e[3] = make_float2(zA.y, dA.y); }
    else           { e[0] = make_float2(zA.y, dA.y); e[1] = make_float2(zB.y, dB.y); e[2] = make_float2(zB.x, dB.x); e[3] = make_float2(zA.x, dA.x); }
    // lane aggregate: zero-state end state of the whole 65-sample chunk
    float2 v = e[0];
#pragma unroll
    for (int m = 1; m < KW_NSUB; ++m) {
        const int j = BACKWARD ? KW_NSUB - 1 - m : m;
        const float* cm = (kw_off(j + 1) - kw_off(j) == 17) ? c.c17 : c.c16;
        const float2 w = kw32_mat(cm, v);
        v = make_float2(w.x + e[m].x, w.y + e[m].y);
    }
    // 2. scan of the chunk end states across lanes
    if (pos == 0) {
        const float2 w = kw32_mat(c.phi[0], s0);
        v.x += w.x; v.y += w.y;
    }
#pragma unroll
    for (int jj = 0; jj < 5; ++jj) {
        const int dd = 1 << jj;
        const float rx = BACKWARD ? __shfl_down_sync(0xffffffffu, v.x, dd) : __shfl_up_sync(0xffffffffu, v.x, dd);
        const float ry = BACKWARD ? __shfl_down_sync(0xffffffffu, v.y, dd) : __shfl_up_sync(0xffffffffu, v.y, dd);
        if (pos >= dd) {
            const float2 w = kw32_mat(c.phi[jj], make_float2(rx, ry));
            v.x += w.x; v.y += w.y;
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
        const float2 w = kw32_mat(cm, sin[m - 1]);
        sin[m] = make_float2(w.x + e[m - 1].x, w.y + e[m - 1].y);
    }
    // 3. homogeneous correction (y = b0 x + z: the state only enters through z): sample of step n of a sub-chunk gets
    //    g[n] . (entry state of that sub-chunk)
    // entry states per pair component: A = (sub-chunk 0, sub-chunk 3), B = (sub-chunk 1, sub-chunk 2)
    const float2 s_c0 = BACKWARD ? sin[3] : sin[0], s_c1 = BACKWARD ? sin[2] : sin[1];
    const float2 s_c2 = BACKWARD ? sin[1] : sin[2], s_c3 = BACKWARD ? sin[0] : sin[3];
    const float2 SXA = make_float2(s_c0.x, s_c3.x), SYA = make_float2(s_c0.y, s_c3.y);
    const float2 SXB = make_float2(s_c1.x, s_c2.x), SYB = make_float2(s_c1.y, s_c2.y);
    if (!BACKWARD) {
        // sub-chunk 0: sample i is step i (r0: step 0, a[k].x: step k + 1); sub-chunks 1 .. 3: slot k is step k
        q.r0 = fmaf(c.g[0][0], s_c0.x, fmaf(c.g[0][1], s_c0.y, q.r0));
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            q.a[k] = __ffma2_rn(c.ga0[k], SXA, __ffma2_rn(c.ga1[k], SYA, q.a[k]));
            q.b[k] = __ffma2_rn(bc2(c.g[k][0]), SXB, __ffma2_rn(bc2(c.g[k][1]), SYB, q.b[k]));
        }
    } else {
        // descending: slot k of every sub-chunk is step 15 - k; sub-chunk 0's sample 0 is step 16
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float2 G0 = bc2(c.g[15 - k][0]), G1 = bc2(c.g[15 - k][1]);
            q.a[k] = __ffma2_rn(G0, SXA, __ffma2_rn(G1, SYA, q.a[k]));
            q.b[k] = __ffma2_rn(G0, SXB, __ffma2_rn(G1, SYB, q.b[k]));
        }
        q.r0 = fmaf(c.g[16][0], s_c0.x, fmaf(c.g[16][1], s_c0.y, q.r0));
    }
}

// odd reflection padding (9 samples) around the frame held at ext positions [9, 2057); the slack of lane 31
// continues the last padded sample (constant input keeps the steady state, see the header)
__device__ __forceinline__ void kw32_odd_pad(Kw32R& q, int lane) {
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < KW_PAD; ++j) kwr(q, j) = 2.f * kwr(q, KW_PAD) - kwr(q, 2 * KW_PAD - j);
    }
    if (lane == 31) {
        constexpr int E = KW_LAST - KW_PAD;
#pragma unroll
        for (int j = 0; j < KW_PAD; ++j) kwr(q, E + 1 + j) = 2.f * kwr(q, E) - kwr(q, E - 1 - j);
#pragma unroll
        for (int i = KW_LAST + 1; i < KW_L; ++i) kwr(q, i) = kwr(q, KW_LAST);
    }
}

__device__ __forceinline__ void kw32_fill_slack(Kw32R& q, int lane) {
    if (lane == 31) {
#pragma unroll
        for (int i = KW_LAST + 1; i < KW_L; ++i) kwr(q, i) = kwr(q, KW_LAST);
    }
}

__device__ __forceinline__ void kw32_zero_pads(Kw32R& q, int lane) {
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < KW_PAD; ++i) kwr(q, i) = 0.f;
    }
    if (lane == 31) {
#pragma unroll
        for (int i = KW_LAST - KW_PAD + 1; i < KW_L; ++i) kwr(q, i) = 0.f;
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
        if (a.hop_inv) {                                                 // warp-uniform
            float m = 0.f;
#pragma unroll
            for (int j = 3 * KW_W / 4 / 4 / 32; j < KW_W / 4 / 32; ++j)    // the frame's last quarter = hop block f
                m = fmaxf(fmaxf(m, fmaxf(fabsf(tmp[j].x), fabsf(tmp[j].y))), fmaxf(fabsf(tmp[j].z), fabsf(tmp[j].w)));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) a.hop_inv[(size_t)ch * a.hop_inv_nb + (f - a.hop_inv_j0)] = hop_row_scale(m);
        }
        if (lane < KW_PAD) stg[KW_STG_SHIFT + lane] = 0.f;
        if (lane < KW_SLACK + KW_PAD + KW_STG_SHIFT) stg[KW_STG_SHIFT + KW_PAD + KW_W + lane] = 0.f;
#pragma unroll
        for (int j = 0; j < KW_W / 4 / 32; ++j) s4[lane + 32 * j] = tmp[j];
        __syncwarp();
        float* sl = stg + KW_STG_SHIFT + KW_L * lane;
        const float* hl = hann_x + KW_L * lane;
        Kw32R r;
        float sumsq = 0.f;
#pragma unroll
        for (int i = 0; i < KW_L; ++i) {
            const float xv = sl[i] * hl[i];
            sumsq = fmaf(xv, xv, sumsq);
            kwr(r, i) = xv;
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
                    kw32_zero_pads(r, lane);
#pragma unroll
                    for (int i = 0; i < KW_L; ++i) sl[i] = kwr(r, i);
                }
            }
            kw32_zero_pads(r, lane);
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < KW_L; ++i) {
                const float fv = sl[i];
                const float w = fmaf(kwr(r, i) - fv, 0.3f, fv);          // f + (s - f) * 0.3 ; 0 at the pads
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
