// waterfall_kernel.cuh -- data side of the spectrogram / waterfall consumers:
//   SpectrogramWaterfall.update + _normalize_spectrum, omega4/panels/spectrogram_waterfall.py:71-121
//     slice = fft_data[lo:hi];  dB = 20 log10(max(slice, 1e-10))                          (:80-85)
//     peak_history.append((max dB, min dB));  with auto gain: current_peak = P95 of the last 20 maxima,
//     current_floor = P5 of the last 20 minima (np.percentile, linear)                    (:88-101)
//     row = clip((dB + gain_adjustment - floor) / (peak - floor), 0, 1), zeros when peak <= floor  (:107-121)
//   SpectrogramPanel.update, omega4/plugins/panels/spectrogram.py:72:  dB = 20 log10(x + 1e-10)  (db_form 1)
//
// The auto gain is a sliding window over per-row extrema, not a recurrence, so every row is independent once
// the extrema are known: pass 1 (one CTA per row) converts the slice and reduces (max, min); pass 2 (one CTA
// per row) gathers the <= 20 most recent extrema -- from this call's rows and the carried state -- evaluates
// the two percentiles and normalises; a one-thread-per-channel kernel then rolls the carried state forward.
// All float32, as the reference computes it.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace o4 {

constexpr int WF_HIST = 20;                       // rows the auto gain looks back over (:96-97)
constexpr int WF_STATE = 1 + 2 * (WF_HIST - 1) + 2;  // count, (max, min) of the last 19 rows oldest first, current peak, floor

struct WaterfallArgs {
    const float* spec;        // [n_ch][n_rows][len]
    int n_ch, n_rows, len, lo, hi;
    int db_form;              // 0: 20 log10(max(x, 1e-10));  1: 20 log10(x + 1e-10)
    int auto_gain;
    float gain_adjustment;
    const float* state_in;    // [n_ch][WF_STATE] or nullptr (fresh: no history, peak 0, floor -80)
    float* state_out;         // [n_ch][WF_STATE] or nullptr
    float* rowstat;           // [n_ch][n_rows][4]: max dB, min dB, current_peak, current_floor after the row
    float* db_out;            // [n_ch][n_rows][hi - lo] or nullptr
    float* norm_out;          // [n_ch][n_rows][hi - lo] or nullptr
};

__device__ __forceinline__ float wf_db(float x, int form) {
    return 20.f * log10f(form ? x + 1e-10f : fmaxf(x, 1e-10f));
}

// np.percentile(v[0..n), q), method 'linear', on an ascending array; numpy's _lerp
__device__ __forceinline__ float wf_percentile(const float* v, int n, float q) {
    const float vi = q * (float)(n - 1);
    int lo = (int)floorf(vi);
    if (lo > n - 2) lo = n - 2;
    if (n < 2) return v[0];
    const float t = vi - (float)lo, a = v[lo], b = v[lo + 1], d = b - a;
    return t >= 0.5f ? b - d * (1.f - t) : a + d * t;
}

__global__ void __launch_bounds__(256)
waterfall_db_kernel(const __grid_constant__ WaterfallArgs a) {
    __shared__ float s_mx[8], s_mn[8];
    const size_t row = blockIdx.x;
    const float* src = a.spec + row * a.len + a.lo;
    const int n = a.hi - a.lo;
    float mx = -CUDART_INF_F, mn = CUDART_INF_F;
    for (int i = threadIdx.x; i < n; i += 256) {
        const float d = wf_db(__ldg(src + i), a.db_form);
        if (a.db_out) a.db_out[row * n + i] = d;
        mx = fmaxf(mx, d); mn = fminf(mn, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 31) == 0) { s_mx[threadIdx.x >> 5] = mx; s_mn[threadIdx.x >> 5] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { mx = fmaxf(mx, s_mx[w]); mn = fminf(mn, s_mn[w]); }
        a.rowstat[row * 4 + 0] = mx;
        a.rowstat[row * 4 + 1] = mn;
    }
}

__global__ void __launch_bounds__(256)
waterfall_norm_kernel(const __grid_constant__ WaterfallArgs a) {
    __shared__ float s_pf[2];
    const size_t row = blockIdx.x;
    const int ch = (int)(row / a.n_rows), k = (int)(row % a.n_rows);
    const int n = a.hi - a.lo;
    if (threadIdx.x == 0) {
        const float* st = a.state_in ? a.state_in + (size_t)ch * WF_STATE : nullptr;
        float peak = st ? st[WF_STATE - 2] : 0.f, floor_ = st ? st[WF_STATE - 1] : -80.f;
        if (a.auto_gain) {
            float mxs[WF_HIST], mns[WF_HIST];
            int cnt = 0;
            const int carried = st ? (int)st[0] : 0;                 // rows of history in the state (<= 19)
            const int from_state = max(0, min(carried, WF_HIST - 1 - k));   // the window reaches that far back
            for (int j = carried - from_state; j < carried; ++j) { mxs[cnt] = st[1 + 2 * j]; mns[cnt] = st[2 + 2 * j]; ++cnt; }
            for (int j = max(0, k - (WF_HIST - 1)); j <= k; ++j) {
                const float* rs = a.rowstat + ((size_t)ch * a.n_rows + j) * 4;
                mxs[cnt] = rs[0]; mns[cnt] = rs[1]; ++cnt;
            }
            for (int i = 1; i < cnt; ++i) {                          // insertion sorts, ascending
                float x = mxs[i]; int j = i - 1;
                while (j >= 0 && mxs[j] > x) { mxs[j + 1] = mxs[j]; --j; }
                mxs[j + 1] = x;
                x = mns[i]; j = i - 1;
                while (j >= 0 && mns[j] > x) { mns[j + 1] = mns[j]; --j; }
                mns[j + 1] = x;
            }
            peak = wf_percentile(mxs, cnt, 0.95f);
            floor_ = wf_percentile(mns, cnt, 0.05f);
        }
        s_pf[0] = peak; s_pf[1] = floor_;
        a.rowstat[row * 4 + 2] = peak;
        a.rowstat[row * 4 + 3] = floor_;
    }
    __syncthreads();
    if (!a.norm_out) return;
    const float peak = s_pf[0], floor_ = s_pf[1], range = peak - floor_;
    const float* src = a.spec + row * a.len + a.lo;
    for (int i = threadIdx.x; i < n; i += 256) {
        float v = 0.f;
        if (range > 0.f) {
            const float d = a.db_out ? a.db_out[row * n + i] : wf_db(__ldg(src + i), a.db_form);
            v = fminf(1.f, fmaxf(0.f, __fdiv_rn((d + a.gain_adjustment) - floor_, range)));
        }
        a.norm_out[row * n + i] = v;
    }
}

// carried state after the call: the (max, min) of the last 19 rows seen so far + the current peak / floor
__global__ void waterfall_state_kernel(const __grid_constant__ WaterfallArgs a) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= a.n_ch || !a.state_out) return;
    const float* st = a.state_in ? a.state_in + (size_t)ch * WF_STATE : nullptr;
    float old[WF_STATE];
    for (int i = 0; i < WF_STATE; ++i) old[i] = st ? st[i] : 0.f;
    if (!st) { old[WF_STATE - 2] = 0.f; old[WF_STATE - 1] = -80.f; }
    const int carried = (int)old[0];
    const int total = min(WF_HIST - 1, carried + a.n_rows);
    float* out = a.state_out + (size_t)ch * WF_STATE;
    const int from_rows = min(a.n_rows, total), from_state = total - from_rows;
    int w = 0;
    for (int j = carried - from_state; j < carried; ++j, ++w) { out[1 + 2 * w] = old[1 + 2 * j]; out[2 + 2 * w] = old[2 + 2 * j]; }
    for (int j = a.n_rows - from_rows; j < a.n_rows; ++j, ++w) {
        const float* rs = a.rowstat + ((size_t)ch * a.n_rows + j) * 4;
        out[1 + 2 * w] = rs[0]; out[2 + 2 * w] = rs[1];
    }
    out[0] = (float)total;
    const float* last = a.rowstat + ((size_t)ch * a.n_rows + a.n_rows - 1) * 4;
    out[WF_STATE - 2] = last[2];
    out[WF_STATE - 1] = last[3];
}

}  // namespace o4
