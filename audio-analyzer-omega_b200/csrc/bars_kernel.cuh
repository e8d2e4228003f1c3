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
// Mapping: the rows of a channel are cut into segments of BARS_SEG hops, one CTA per (channel,
// segment); a CTA walks its hops in order, 8 hops per iteration (one warp per hop does the row work:
// selection of the two order statistics np.percentile interpolates, scaling, band means), then the
// only sequential step -- the first-order smoothing recurrence along the hops -- is applied by one
// thread per band over those 8 rows.  Segment 0 starts from the carried state; every later segment
// first replays the BARS_WARM rows before it, which reproduces the recurrence state to
// 0.85^128 = 1e-9 (smoothing factors are <= 0.85), far below float32 resolution of the [0, 1] bars.
// All fp32.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace o4 {

constexpr int BARS_WARPS = 8;
constexpr int BARS_SEG = 512;        // hops per CTA (BarsArgs::seg: 512, or 2048 when there are plenty of channels)
constexpr int BARS_WARM = 128;       // rows replayed before a segment to rebuild the smoothing state
constexpr int BARS_MAXV = 32;        // spectrum values per lane held in registers (T <= 1024); kernel templated on 16 / 32

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
    float* state;              // [n_ch][1 + n_valid]: has_prev flag, previous bars; or nullptr (written by the last segment)
    const float* state_in;     // what segment 0 reads: `state`, or a copy of it when several segments run side by side
    int fresh;                 // ignore state contents on entry
    float* bars_out;           // [n_ch][n_hops][n_valid]
    float* peaks_out;          // same shape or nullptr (band values before smoothing)
    int seg;                   // hops per CTA
};

template <int N>
__device__ __forceinline__ void bars_bitonic_desc(float* x) {       // x[0..N) in registers, descending on exit
#pragma unroll
    for (int k = 2; k <= N; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const bool desc = ((i & k) == 0);
                    const float hi = fmaxf(x[i], x[l]), lo = fminf(x[i], x[l]);
                    x[i] = desc ? hi : lo;
                    x[l] = desc ? lo : hi;
                }
            }
}

template <int NV>
__global__ void __launch_bounds__(BARS_WARPS * 32, (NV <= 16 ? 4 : 2))
bars_kernel(const __grid_constant__ BarsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int RSTR = 32 * (NV + 1);                        // row buffer: T values, or 32 x 33 candidates
    float* rows = reinterpret_cast<float*>(smem_raw);                 // [WARPS][RSTR] scaled spectrum
    float* raw = rows + BARS_WARPS * RSTR;                            // [WARPS][n_valid] unsmoothed bars
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_seg = (a.n_hops + a.seg - 1) / a.seg;
    const int ch = blockIdx.x / n_seg, seg = blockIdx.x % n_seg;
    const int out_beg = seg * a.seg;                               // first hop this CTA writes
    const int out_end = min(a.n_hops, out_beg + a.seg);
    const int beg = (seg == 0) ? 0 : max(0, out_beg - BARS_WARM);     // first hop it processes
    float* row = rows + warp * RSTR;
    float* rw = raw + warp * a.n_valid;
    const int K = a.T - a.p_lo;                                       // sorted[p_lo] is the K-th largest
    constexpr int nv = NV;                                            // values per lane (T <= 32 NV)

    // smoothing state of the bands this thread owns (thread b handles bands b, b + 256, ...)
    constexpr int MAXB = 8;                                           // up to 2048 bands
    float prev[MAXB];
    bool has_prev = false;
#pragma unroll
    for (int i = 0; i < MAXB; ++i) prev[i] = 0.f;
    if (seg == 0 && a.state && !a.fresh) {
        const float* st = (a.state_in ? a.state_in : a.state) + (size_t)ch * (1 + a.n_valid);
        has_prev = st[0] != 0.f;
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
            const int b = threadIdx.x + i * BARS_WARPS * 32;
            if (b < a.n_valid) prev[i] = st[1 + b];
        }
    }
    // a warm-up that starts at hop 0 of a fresh channel reproduces the sequential run exactly; one that
    // starts later begins from "no previous frame", which the 128-row replay forgets
    // the next iteration's row is fetched while the current one is processed
    float nxt[NV];
    {
        const int hop = beg + warp;
        const float* src = a.spec + ((size_t)ch * a.n_hops + (hop < out_end ? hop : beg)) * a.T;
#pragma unroll
        for (int i = 0; i < NV; ++i) { const int idx = lane + 32 * i; nxt[i] = idx < a.T ? __ldg(src + idx) : -CUDART_INF_F; }
    }
    for (int h0 = beg; h0 < out_end; h0 += BARS_WARPS) {
        const int hop = h0 + warp;
        float v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = nxt[i];
        if (hop + BARS_WARPS < out_end) {
            const float* src = a.spec + ((size_t)ch * a.n_hops + hop + BARS_WARPS) * a.T;
#pragma unroll
            for (int i = 0; i < NV; ++i) { const int idx = lane + 32 * i; nxt[i] = idx < a.T ? __ldg(src + idx) : -CUDART_INF_F; }
        }
        if (hop < out_end) {
            float mx = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) if (lane + 32 * i < a.T) mx = fmaxf(mx, v[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float ref = 0.f;
            bool do_scale = false;
            if (mx > 0.f) {
                // K-th and (K-1)-th largest (K ~ 2 % of T) by K extractions of the warp maximum.  Values
                // are non-negative, so their float bits order like unsigned integers: one REDUX
                // (__reduce_max_sync) finds the maximum, a ballot its owner.  Each lane's candidates are
                // sorted in registers (bitonic network over its 16 or 32 values) and parked as a descending
                // list in its slice of the warp's row buffer, so the owner's next candidate is one LDS away.
                // Slots past the end of the row are padded with -1 for the sort and enter the selection as bit
                // pattern 0 (the bits of -1.f would beat every real value in the unsigned comparison).
                float srt[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i) srt[i] = (lane + 32 * i < a.T) ? v[i] : -1.f;
                bars_bitonic_desc<NV>(srt);
                float* cand = row + lane * (NV + 1);                  // odd lane stride: conflict free
#pragma unroll
                for (int i = 0; i < NV; ++i) cand[i] = srt[i];
                int ptr = 0;
                unsigned lbits = __float_as_uint(fmaxf(cand[0], 0.f));
                float kth = 0.f, kth1 = 0.f;                          // sorted[p_lo], sorted[p_lo + 1]
                for (int it = 1; it <= K; ++it) {
                    const unsigned wb = __reduce_max_sync(0xffffffffu, lbits);
                    const int wl = __ffs(__ballot_sync(0xffffffffu, lbits == wb)) - 1;
                    if (it == K - 1) kth1 = __uint_as_float(wb);
                    if (it == K) kth = __uint_as_float(wb);
                    if (lane == wl) { ++ptr; lbits = ptr < nv ? __float_as_uint(fmaxf(cand[ptr], 0.f)) : 0u; }
                }
                __syncwarp();
                if (K == 1) kth1 = kth;
                // numpy _lerp: a + (b - a) t, or b - (b - a)(1 - t) for t >= 0.5
                const float d = kth1 - kth;
                ref = (a.p_frac >= 0.5f) ? kth1 - d * (1.f - a.p_frac) : kth + d * a.p_frac;
                do_scale = ref > 0.f;
            }
            // scale, compensate; optional max normalisation
            float mx2 = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int idx = lane + 32 * i;
                if (idx < a.T) {
                    float x = v[i];
                    if (do_scale) x = __fdiv_rn(x, ref) * a.scale;
                    if (a.gain) x *= __ldg(a.gain + idx);
                    v[i] = x;
                    mx2 = fmaxf(mx2, x);
                }
            }
            if (a.normalize_max) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
            }
            const bool norm = a.normalize_max && mx2 > 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int idx = lane + 32 * i;
                if (idx < a.T) row[idx] = norm ? __fdiv_rn(v[i], mx2) : v[i];
            }
            __syncwarp();
            for (int b = lane; b < a.n_valid; b += 32) {
                const int s = __ldg(a.bands + 2 * b), e = __ldg(a.bands + 2 * b + 1);
                float acc = 0.f;
                for (int i = s; i < e; ++i) acc += row[i];
                float x = (e > s) ? __fdiv_rn(acc, (float)(e - s)) : 0.f;
                if (x > 0.f) x = fminf(1.f, sqrtf(x));
                rw[b] = x;
            }
        }
        __syncthreads();
        // smoothing recurrence over the rows of this iteration, one thread per band
        const int nrow = min(BARS_WARPS, out_end - h0);
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
            const int b = threadIdx.x + i * BARS_WARPS * 32;
            if (b >= a.n_valid) break;
            float p = prev[i];
            bool hp = has_prev;
            const float sf = a.sf ? a.sf[b] : 0.f, sfc = a.sf ? a.sfc[b] : 1.f;
            for (int r = 0; r < nrow; ++r) {
                const float x = raw[r * a.n_valid + b];
                float y = x;
                if (a.sf && hp) y = __fadd_rn(__fmul_rn(p, sf), __fmul_rn(x, sfc));
                if (h0 + r >= out_beg) {
                    const size_t o = ((size_t)ch * a.n_hops + h0 + r) * a.n_valid + b;
                    a.bars_out[o] = y;
                    if (a.peaks_out) a.peaks_out[o] = x;
                }
                p = y; hp = true;
            }
            prev[i] = p;
        }
        has_prev = true;
        __syncthreads();
    }
    if (a.state && seg == n_seg - 1) {                                // the last segment carries the state out
        float* st = a.state + (size_t)ch * (1 + a.n_valid);
        if (threadIdx.x == 0) st[0] = 1.f;
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
            const int b = threadIdx.x + i * BARS_WARPS * 32;
            if (b < a.n_valid) st[1 + b] = prev[i];
        }
    }
}

inline size_t bars_smem_bytes(int T, int n_valid) {
    const int nv = T <= 512 ? 16 : 32;
    return (size_t)BARS_WARPS * (32 * (nv + 1) + n_valid) * sizeof(float) + 16;
}

}  // namespace o4
