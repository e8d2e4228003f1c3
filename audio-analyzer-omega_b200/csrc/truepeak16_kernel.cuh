// truepeak16_kernel.cuh -- the batch path's 4x true peak with the transforms evaluated in half precision, TWO frame
// pairs (four frames) per transform.
//
// Same contract, same formulation as truepeak_kernel.cuh (scipy.signal.resample(x, 4 len(x)) -> max |.| -> 20 log10,
// omega4/panels/professional_meters.py:283-299): frames a, b packed as z = a + i b, one forward transform, then for the
// phases p = 1, 2, 3 the "inverse" transform of Z .* R_p, of which only max |Re|, max |Im| are needed.  What changes:
//   * phase 0 (the samples themselves: the sample peak, and the scale that brings every frame to unit peak) stays
//     float32 and exact;
//   * the forward transform and the three delayed phases run with HFMA2 / HADD2 on (pair A, pair B) half2 lanes: every
//     butterfly instruction, every shared-memory exchange and every barrier serves two frame pairs; the spectrum is
//     kept as half (re, im) of both pairs in shared memory, stage twiddles and the delay factors R_p are packed half.
// What it buys (tools/micro/f32x2_pipes.cu, profiles/r02l_f32x2_pipes.txt): HADD2 / HFMA2 / HMUL2 issue at 1.9 warp
// instructions per clock and SM against 3.7 for FADD / FFMA -- a half2 instruction costs the FMA pipe two slots, so the
// arithmetic rate per transform is the float32 one; the gain is in everything else (issue slots, shared-memory
// wavefronts, barriers, registers): 102.6 ms (float32) -> 77.5 (delayed phases in half, r02k) -> 65 ms (all four
// transforms, r02l) per 11.52 M frames.
// Accuracy: a delayed phase only matters where it exceeds the sample peak, and then as a maximum of values of the
// frame's own size; half keeps 11 significant bits through 11 butterfly levels, which leaves the peak within
// ~2e-3 dB typically and 0.025 dB at worst (3600 emulated frames) of the float64 reference (north star: 0.05 dBTP; numpy emulation of this
// arithmetic over sines, noise, clipped noise, square waves, impulses, random walks: tests/tools/truepeak16_numerics.py
// -- the forward transform's rounding adds little to what the three inverse transforms carry: max 0.023 instead of
// 0.022, p99 0.013 / 0.011 over 600 frames; on the GPU: the golden / stress tests, <= 0.021).  The float32 kernel
// stays the one the explicit-frame entry points (omega4_meter_frames, the streaming shim's calculate_true_peak) run,
// and OMEGA4_FLAG_EXACT_TRUE_PEAK / OMEGA4_TP_F32=1 select it for the batch path as well.
#pragma once
#include <cuda_fp16.h>
#include "truepeak_kernel.cuh"

namespace o4 {

struct c2h { __half2 x, y; };        // one complex value of two transforms: (re_A, re_B), (im_A, im_B)

__device__ __forceinline__ c2h hc_add(c2h a, c2h b) { c2h r; r.x = __hadd2(a.x, b.x); r.y = __hadd2(a.y, b.y); return r; }
__device__ __forceinline__ c2h hc_sub(c2h a, c2h b) { c2h r; r.x = __hsub2(a.x, b.x); r.y = __hsub2(a.y, b.y); return r; }
__device__ __forceinline__ c2h hc_mul_mi(c2h a) { c2h r; r.x = a.y; r.y = __hneg2(a.x); return r; }       // * -i
// multiply by a twiddle (re, im) packed in one half2, the same for both transforms
__device__ __forceinline__ c2h hc_mul(c2h a, __half2 w) {
    const __half2 wr = __low2half2(w), wi = __high2half2(w);
    c2h r;
    r.x = __hfma2(a.x, wr, __hneg2(__hmul2(a.y, wi)));
    r.y = __hfma2(a.x, wi, __hmul2(a.y, wr));
    return r;
}
__device__ __forceinline__ c2h hc_mul_c(c2h a, float cr, float ci) { return hc_mul(a, __floats2half2_rn(cr, ci)); }
// * (r, -r), * (-r, -r) with r = sqrt(1/2)
__device__ __forceinline__ c2h hc_rot_m45(c2h a) {
    const __half2 r = __float2half2_rn(0.70710678118654752440f);
    c2h o; o.x = __hmul2(__hadd2(a.x, a.y), r); o.y = __hmul2(__hsub2(a.y, a.x), r); return o;
}
__device__ __forceinline__ c2h hc_rot_m135(c2h a) {
    const __half2 r = __float2half2_rn(0.70710678118654752440f);
    c2h o; o.x = __hmul2(__hsub2(a.y, a.x), r); o.y = __hneg2(__hmul2(__hadd2(a.x, a.y), r)); return o;
}

__device__ __forceinline__ void hbf4(c2h& x0, c2h& x1, c2h& x2, c2h& x3) {
    const c2h a = hc_add(x0, x2), b = hc_sub(x0, x2), c = hc_add(x1, x3), d = hc_mul_mi(hc_sub(x1, x3));
    x0 = hc_add(a, c); x1 = hc_add(b, d); x2 = hc_sub(a, c); x3 = hc_sub(b, d);
}
__device__ __forceinline__ void hbf8(c2h* v) {
    c2h e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    c2h o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    hbf4(e0, e1, e2, e3);
    hbf4(o0, o1, o2, o3);
    o1 = hc_rot_m45(o1);
    o2 = hc_mul_mi(o2);
    o3 = hc_rot_m135(o3);
    v[0] = hc_add(e0, o0); v[4] = hc_sub(e0, o0);
    v[1] = hc_add(e1, o1); v[5] = hc_sub(e1, o1);
    v[2] = hc_add(e2, o2); v[6] = hc_sub(e2, o2);
    v[3] = hc_add(e3, o3); v[7] = hc_sub(e3, o3);
}
// forward radix-16 butterfly, in place, natural order out (bf16pt of fft_core.cuh on half2 lanes)
__device__ __forceinline__ void hbf16(c2h* v) {
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
#pragma unroll
    for (int a = 0; a < 4; ++a) hbf4(v[a], v[a + 4], v[a + 8], v[a + 12]);
    v[1 + 4] = hc_mul_c(v[1 + 4], c1, -s1);
    v[1 + 8] = hc_rot_m45(v[1 + 8]);
    v[1 + 12] = hc_mul_c(v[1 + 12], s1, -c1);
    v[2 + 4] = hc_rot_m45(v[2 + 4]);
    v[2 + 8] = hc_mul_mi(v[2 + 8]);
    v[2 + 12] = hc_rot_m135(v[2 + 12]);
    v[3 + 4] = hc_mul_c(v[3 + 4], s1, -c1);
    v[3 + 8] = hc_rot_m135(v[3 + 8]);
    v[3 + 12] = hc_mul_c(v[3 + 12], -c1, s1);
#pragma unroll
    for (int c = 0; c < 4; ++c) hbf4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) { const c2h tmp = v[4 * c + d]; v[4 * c + d] = v[4 * d + c]; v[4 * d + c] = tmp; }
}
// stage twiddles as packed half (re, im): 15 registers per stage, no conversions in the loop
struct TwH15 { __half2 w[15]; };
struct LocalTwH { TwH15 s1, s2; };
__device__ __forceinline__ void h_apply_twiddles(c2h* v, const TwH15& t) {
#pragma unroll
    for (int k = 1; k < 16; ++k) v[k] = hc_mul(v[k], t.w[k - 1]);
}
template <int LOG2M>
__device__ __forceinline__ void pack_local_twiddles(const LocalTwFull<LOG2M>& a, LocalTwH& o) {
#pragma unroll
    for (int k = 0; k < 15; ++k) {
        o.s1.w[k] = __floats2half2_rn(a.s1.w[k].x, a.s1.w[k].y);
        o.s2.w[k] = __floats2half2_rn(a.s2.w[k].x, a.s2.w[k].y);
    }
}

// fft_forward_local<LOG2M, true> on half2 lanes: on entry v[16] = z[t + j M/16] of both transforms, on exit the last
// stage's outputs are in v[] (nothing is written to a Z buffer).  One group barrier inside (after the stage-1 store);
// the caller must put another one before the next store into X.
template <int LOG2M, class TW>
__device__ __forceinline__ void fft_local_h2(c2h* v, c2h* X, const TW& st, int t, int g) {
    constexpr int M = 1 << LOG2M, TPF = M / 16, G2 = M / 256, S = 17 * G2;
    static_assert(G2 == 8, "written for 2048 complex points");
    hbf16(v);
    h_apply_twiddles(v, st.s1);
#pragma unroll
    for (int k = 0; k < 16; ++k) X[k * S + t] = v[k];
    group_sync<TPF>(g);
    const int q = t / G2, p = t % G2;
    c2h* Xq = X + q * S;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = Xq[p + G2 * j];
    hbf16(v);
    h_apply_twiddles(v, st.s2);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) Xq[17 * p + k] = v[k];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16 / G2; ++i) {
        const int k = p + G2 * i;
#pragma unroll
        for (int pp = 0; pp < G2; ++pp) v[i * G2 + pp] = Xq[17 * pp + k];
        hbf8(v + i * 8);
    }
}

// The same transform with the natural-order spectrum of both pairs written to Zc[zaddr(k)] (one c2h = 8 bytes per
// bin).  Two group barriers (after the stage-1 store, after the final store).
template <int LOG2M, class TW>
__device__ __forceinline__ void fft_forward_h2_z(c2h* v, c2h* X, c2h* Zc, const TW& st, int t, int g) {
    constexpr int M = 1 << LOG2M, TPF = M / 16, G2 = M / 256, S = 17 * G2;
    static_assert(G2 == 8, "written for 2048 complex points");
    hbf16(v);
    h_apply_twiddles(v, st.s1);
#pragma unroll
    for (int k = 0; k < 16; ++k) X[k * S + t] = v[k];
    group_sync<TPF>(g);
    const int q = t / G2, p = t % G2;
    const int zq = zaddr<LOG2M>(q + 16 * p);
    c2h* Xq = X + q * S;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = Xq[p + G2 * j];
    hbf16(v);
    h_apply_twiddles(v, st.s2);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) Xq[17 * p + k] = v[k];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16 / G2; ++i) {
        const int k = p + G2 * i;
#pragma unroll
        for (int pp = 0; pp < G2; ++pp) v[i * G2 + pp] = Xq[17 * pp + k];
        hbf8(v + i * 8);
#pragma unroll
        for (int pq = 0; pq < G2; ++pq) Zc[zq + 16 * G2 * i + 256 * pq] = v[i * G2 + pq];
    }
    group_sync<TPF>(g);
}

// One unit = two frame pairs (four frames) per 128-thread group: sample peaks (float32, = phase 0), scale to unit
// peak, one forward and three "inverse" half2 transforms, maxima.
template <int LOG2M>
__global__ void __launch_bounds__(256, 2)
truepeak16_kernel(const __grid_constant__ TruePeakArgs a) {
    using S = FftShape<LOG2M>;
    constexpr int M = S::M, TPF = S::TPF, CONC = S::CONC, BUF = S::BUF;
    constexpr int WARPS = TPF / 32;
    static_assert(TPF >= 32 && S::NT == 256 && LOG2M == 11, "true-peak kernel: 2048 complex points, whole warps per transform");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* bufs = reinterpret_cast<float2*>(smem_raw);                            // [CONC][BUF + M] 8-byte elements
    float4* red_all = reinterpret_cast<float4*>(bufs + (size_t)CONC * (BUF + M));  // [CONC][2][WARPS] maxima (A.a, A.b, B.a, B.b)

    const int tid = threadIdx.x;
    const int g = tid / TPF;
    const int t = tid % TPF;
    c2h* Xh = reinterpret_cast<c2h*>(bufs + (size_t)g * (BUF + M));
    c2h* Zc = Xh + BUF;                                                            // [M] spectrum of pair A, pair B
    float4* red = red_all + g * 2 * WARPS;

    const int pairs_per_cta = a.rounds * CONC;
    const int n_pairs = (a.n_frames + 1) >> 1;
    const int tiles_per_ch = (n_pairs + pairs_per_cta - 1) / pairs_per_cta;
    const int ch = blockIdx.x / tiles_per_ch;
    const int tile = blockIdx.x % tiles_per_ch;

    LocalTwH st;
    {
        LocalTw<LOG2M> st4;
        LocalTwFull<LOG2M> stf;
        load_local_twiddles<LOG2M>(st4, a.twM, t);
        expand_local_twiddles<LOG2M>(st4, stf);
        pack_local_twiddles<LOG2M>(stf, st);
    }
    const float inv_m = 1.0f / (float)M;
    auto group_max4 = [&](float4 v, int slot) -> float4 {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v.x = fmaxf(v.x, __shfl_xor_sync(0xffffffffu, v.x, o));
            v.y = fmaxf(v.y, __shfl_xor_sync(0xffffffffu, v.y, o));
            v.z = fmaxf(v.z, __shfl_xor_sync(0xffffffffu, v.z, o));
            v.w = fmaxf(v.w, __shfl_xor_sync(0xffffffffu, v.w, o));
        }
        if ((t & 31) == 0) red[slot * WARPS + (t >> 5)] = v;
        group_sync<TPF>(g);
        float4 m = red[slot * WARPS];
#pragma unroll
        for (int w = 1; w < WARPS; ++w) {
            const float4 o = red[slot * WARPS + w];
            m.x = fmaxf(m.x, o.x); m.y = fmaxf(m.y, o.y); m.z = fmaxf(m.z, o.z); m.w = fmaxf(m.w, o.w);
        }
        return m;
    };

    const int zt = zaddr<LOG2M>(t);
    // windowed samples of a pair: v[j] = (a[t + j TPF], b[t + j TPF])
    auto load_pair = [&](int pi, float2 (&v)[16], bool& act_a, bool& act_b) {
        const int fa = 2 * pi, fb = fa + 1;
        act_a = (fa < a.n_frames) && (fa >= a.first_frame);
        act_b = (fb < a.n_frames) && (fb >= a.first_frame);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = make_float2(0.f, 0.f);
        if (!(act_a || act_b)) return;
        const float* pa = reinterpret_cast<const float*>(a.x) + (long long)ch * a.ch_stride + a.frame_off0 + (long long)fa * a.frame_stride + t;
        const float* pb = pa + a.frame_stride;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (act_a) v[j].x = __ldg(pa + j * TPF);
            if (act_b) v[j].y = __ldg(pb + j * TPF);
        }
        if (a.window) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { const float w = __ldg(a.window + t + j * TPF); v[j].x *= w; v[j].y *= w; }
        }
    };

    for (int u = 0; 2 * u < a.rounds; ++u) {
        const int pa_i = tile * pairs_per_cta + (2 * u) * CONC + g;
        const int pb_i = tile * pairs_per_cta + (2 * u + 1) * CONC + g;
        const bool has_b = (2 * u + 1) < a.rounds;
        bool aa, ab, ba = false, bb = false;
        float4 pk4 = make_float4(0.f, 0.f, 0.f, 0.f);
        c2h h[16];
        {
            float2 va[16], vb[16];
            load_pair(pa_i, va, aa, ab);
            if (has_b) load_pair(pb_i, vb, ba, bb);
            else {
#pragma unroll
                for (int j = 0; j < 16; ++j) vb[j] = make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                pk4.x = fmaxf(pk4.x, fabsf(va[j].x)); pk4.y = fmaxf(pk4.y, fabsf(va[j].y));
                pk4.z = fmaxf(pk4.z, fabsf(vb[j].x)); pk4.w = fmaxf(pk4.w, fabsf(vb[j].y));
            }
            pk4 = group_max4(pk4, 0);                   // the four sample peaks = phase 0, exact
            const float s0 = pk4.x > 1e-20f ? 1.f / pk4.x : 0.f, s1 = pk4.y > 1e-20f ? 1.f / pk4.y : 0.f;
            const float s2 = pk4.z > 1e-20f ? 1.f / pk4.z : 0.f, s3 = pk4.w > 1e-20f ? 1.f / pk4.w : 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {              // every frame to unit peak, (A, B) on the half2 lanes
                h[j].x = __floats2half2_rn(va[j].x * s0, vb[j].x * s2);
                h[j].y = __floats2half2_rn(va[j].y * s1, vb[j].y * s3);
            }
        }
        __half2 mre = __float2half2_rn(0.f), mim = mre;            // (A, B) maxima of |Re| (frames a) and |Im| (frames b)
        if (aa || ab || ba || bb) {                                // group-uniform
            fft_forward_h2_z<LOG2M>(h, Xh, Zc, st, t, g);
#pragma unroll 1
            for (int p = 1; p <= 3; ++p) {
                const unsigned* rot_p = a.rot_h + (p - 1) * M + t;      // delay factors as packed half, 24 KB, L1 resident
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const unsigned rku = __ldg(rot_p + j * TPF);
                    const __half2 rkh = *reinterpret_cast<const __half2*>(&rku);
                    const c2h s = hc_mul(Zc[zt + j * TPF], rkh);
                    h[j].x = s.x; h[j].y = __hneg2(s.y);          // conj: inverse transform by the forward kernel
                }
                fft_local_h2<LOG2M>(h, Xh, st, t, g);
#pragma unroll
                for (int j = 0; j < 16; ++j) { mre = __hmax2(mre, __habs2(h[j].x)); mim = __hmax2(mim, __habs2(h[j].y)); }
                group_sync<TPF>(g);      // every thread is past this transform's X reads before the next stage-1 store
            }
        }
        const float2 fre = __half22float2(mre), fim = __half22float2(mim);
        const float4 pko = group_max4(make_float4(fre.x, fim.x, fre.y, fim.y), 1);     // (A.a, A.b, B.a, B.b)
        if (t == 0) {
            const float pk[4] = {pk4.x, pk4.y, pk4.z, pk4.w};
            const float po[4] = {pko.x, pko.y, pko.z, pko.w};
            const bool act[4] = {aa, ab, ba, bb};
            const int fr[4] = {2 * pa_i, 2 * pa_i + 1, 2 * pb_i, 2 * pb_i + 1};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!act[i]) continue;
                const double peak = (double)pk[i] * (double)fmaxf(1.f, po[i] * inv_m);     // normalised peaks (phase 0 = 1)
                a.tp_out[(size_t)ch * a.n_frames + fr[i]] = (peak < 1e-10) ? -100.0 : 20.0 * log10(peak);
            }
        }
        group_sync<TPF>(g);          // red[] and Zc are reused by the next unit
    }
}

template <int LOG2M>
inline size_t truepeak16_smem_bytes() {
    using S = FftShape<LOG2M>;
    return (size_t)S::CONC * (S::BUF + S::M) * sizeof(float2) + (size_t)S::CONC * 2 * (S::TPF / 32) * sizeof(float4) + 16;
}

}  // namespace o4
