// bars_kernel.cuh -- the application's spectrum post-processing between the combined spectrum and
// the drawable `band_values`:
//   ProfessionalLiveAudioAnalyzer.process_audio_spectrum, omega4_main.py:992-1056
//     P98 = np.percentile(spectrum, 98); spectrum = spectrum / P98 * 0.8        (:992-998)
//     spectrum *= frequency-compensation gains (apply_frequency_compensation, :855-926 -> table)
//     optional spectrum /= max(spectrum)                                        (:1005-1006)
//     bar = clamp(sqrt(mean(spectrum[s:e])), 0, 1) over the mel band table, stopping at the first
//           band that reaches past the spectrum                                 (:1012-1032)
//     bar = prev * sf + bar * (1 - sf), sf per band, prev = last frame's bars   (:1041-1056)
//
// Mapping: one CTA per channel walks its hops in order, 8 hops per iteration (one warp per hop does
// the row work: selection of the two order statistics np.percentile interpolates, scaling, band
// means); the only sequential step -- the first-order smoothing recurrence along the hops -- is
// then applied by one thread per band over those 8 rows.  All fp32.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace o4 {

constexpr int BARS_WARPS = 8;

struct BarsArgs {
    const float* spec;         // [n_ch][n_hops][T]
    int T;
    int n_ch, n_hops;
    int n_valid;               // bands with end <= T
    const int* bands;          // [n_valid][2]
    const float* gain;         // [T] or nullptr
    const float* sf;           // [n_valid] smoothing factor, or nullptr (smoothing off)
    const float* sfc;          // [n_valid] float32(1 - sf)
    int p_lo;                  // np.percentile: sorted[p_lo] + frac * (sorted[p_lo + 1] - sorted[p_lo])
    float p_frac;
    float scale;               // 0.8
    int normalize_max;
    float* state;              // [n_ch][1 + n_valid]: has_prev flag, previous bars; or nullptr
    int fresh;                 // ignore state contents on entry
    float* bars_out;           // [n_ch][n_hops][n_valid]
    float* peaks_out;          // same shape or nullptr (band values before smoothing)
};

__global__ void __launch_bounds__(BARS_WARPS * 32)
bars_kernel(const __grid_constant__ BarsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* rows = reinterpret_cast<float*>(smem_raw);                 // [WARPS][T] scaled spectrum
    float* tmp = rows + BARS_WARPS * a.T;                             // [WARPS][T] selection scratch
    float* raw = tmp + BARS_WARPS * a.T;                              // [WARPS][n_valid] unsmoothed bars
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = blockIdx.x;
    float* row = rows + warp * a.T;
    float* sel = tmp + warp * a.T;
    float* rw = raw + warp * a.n_valid;
    const int K = a.T - a.p_lo;                                       // sorted[p_lo] is the K-th largest

    // smoothing state of the bands this thread owns (thread b handles bands b, b + 256, ...)
    constexpr int MAXB = 8;                                           // up to 2048 bands
    float prev[MAXB];
    bool has_prev = false;
    if (a.state && !a.fresh) {
        const float* st = a.state + (size_t)ch * (1 + a.n_valid);
        has_prev = st[0] != 0.f;
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
            const int b = threadIdx.x + i * BARS_WARPS * 32;
            prev[i] = (b < a.n_valid) ? st[1 + b] : 0.f;
        }
    } else {
#pragma unroll
        for (int i = 0; i < MAXB; ++i) prev[i] = 0.f;
    }

    for (int h0 = 0; h0 < a.n_hops; h0 += BARS_WARPS) {
        const int hop = h0 + warp;
        if (hop < a.n_hops) {
            const float* src = a.spec + ((size_t)ch * a.n_hops + hop) * a.T;
            float mx = 0.f;
            for (int i = lane; i < a.T; i += 32) {
                const float v = src[i];
                row[i] = v; sel[i] = v;
                mx = fmaxf(mx, v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            __syncwarp();
            float ref = 0.f;
            bool do_scale = false;
            if (mx > 0.f) {
                // K-th and (K-1)-th largest by repeated extraction of the maximum (K ~ 2 % of T)
                float lmax = -CUDART_INF_F; int lidx = -1;
                for (int i = lane; i < a.T; i += 32) { const float v = sel[i]; if (v > lmax) { lmax = v; lidx = i; } }
                float kth = 0.f, kth1 = 0.f;                          // sorted[p_lo], sorted[p_lo + 1]
                for (int it = 1; it <= K; ++it) {
                    float wv = lmax; int wl = lane;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
                        const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
                        if (ov > wv || (ov == wv && ol < wl)) { wv = ov; wl = ol; }
                    }
                    if (it == K - 1) kth1 = wv;
                    if (it == K) kth = wv;
                    if (lane == wl) {                                 // winner removes its element, rescans its slice
                        sel[lidx] = -CUDART_INF_F;
                        lmax = -CUDART_INF_F; lidx = -1;
                        for (int i = lane; i < a.T; i += 32) { const float v = sel[i]; if (v > lmax) { lmax = v; lidx = i; } }
                    }
                }
                if (K == 1) kth1 = kth;
                // numpy _lerp: a + (b - a) t, or b - (b - a)(1 - t) for t >= 0.5
                const float d = kth1 - kth;
                ref = (a.p_frac >= 0.5f) ? kth1 - d * (1.f - a.p_frac) : kth + d * a.p_frac;
                do_scale = ref > 0.f;
            }
            // scale, compensate; optional max normalisation
            float mx2 = 0.f;
            for (int i = lane; i < a.T; i += 32) {
                float v = row[i];
                if (do_scale) v = __fdiv_rn(v, ref) * a.scale;
                if (a.gain) v *= a.gain[i];
                row[i] = v;
                mx2 = fmaxf(mx2, v);
            }
            if (a.normalize_max) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
                __syncwarp();
                if (mx2 > 0.f)
                    for (int i = lane; i < a.T; i += 32) row[i] = __fdiv_rn(row[i], mx2);
            }
            __syncwarp();
            for (int b = lane; b < a.n_valid; b += 32) {
                const int s = a.bands[2 * b], e = a.bands[2 * b + 1];
                float acc = 0.f;
                for (int i = s; i < e; ++i) acc += row[i];
                float v = (e > s) ? __fdiv_rn(acc, (float)(e - s)) : 0.f;
                if (v > 0.f) v = fminf(1.f, sqrtf(v));
                rw[b] = v;
            }
        }
        __syncthreads();
        // smoothing recurrence over the rows of this iteration, one thread per band
        const int nrow = min(BARS_WARPS, a.n_hops - h0);
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
            const int b = threadIdx.x + i * BARS_WARPS * 32;
            if (b >= a.n_valid) break;
            float p = prev[i];
            bool hp = has_prev;
            const float sf = a.sf ? a.sf[b] : 0.f, sfc = a.sf ? a.sfc[b] : 1.f;
            for (int r = 0; r < nrow; ++r) {
                const float v = raw[r * a.n_valid + b];
                float y = v;
                if (a.sf && hp) y = __fadd_rn(__fmul_rn(p, sf), __fmul_rn(v, sfc));
                const size_t o = ((size_t)ch * a.n_hops + h0 + r) * a.n_valid + b;
                a.bars_out[o] = y;
                if (a.peaks_out) a.peaks_out[o] = v;
                p = y; hp = true;
            }
            prev[i] = p;
        }
        has_prev = true;
        __syncthreads();
    }
    if (a.state) {
        float* st = a.state + (size_t)ch * (1 + a.n_valid);
        if (threadIdx.x == 0) st[0] = (has_prev || a.n_hops > 0) ? 1.f : 0.f;
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
            const int b = threadIdx.x + i * BARS_WARPS * 32;
            if (b < a.n_valid) st[1 + b] = prev[i];
        }
    }
}

inline size_t bars_smem_bytes(int T, int n_valid) {
    return (size_t)BARS_WARPS * (2 * T + n_valid) * sizeof(float) + 16;
}

}  // namespace o4
