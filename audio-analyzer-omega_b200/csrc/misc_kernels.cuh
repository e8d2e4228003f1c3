// misc_kernels.cuh -- small kernels around the hot path: general combine (overlapping ranges /
// caller-supplied magnitudes), mel band mapping, synthetic audio generator.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace o4 {

// ---------------------------------------------------------------------------------------------
// General combine: MultiResolutionFFT.combine_results_optimized (multi_resolution_fft.py:359-395)
// for any set of present resolutions, including overlapping freq_ranges.  CSR over target bins.
// ---------------------------------------------------------------------------------------------
struct CombineArgs {
    const float* mag[8];       // per resolution [n_rows][bins[r]] or nullptr (absent)
    int bins[8];               // N_r/2+1
    int first_frame[8];        // rows with (row % n_hops) < first_frame[r] have resolution r absent
    float weight[8];
    int n_hops;                // rows per channel (for the readiness schedule); 0 = always present
    int n_rows;
    int T;
    const int* csr_ptr;        // [T+1]
    const int* csr_res;        // [nnz]
    const int* csr_lo;         // [nnz]
    const float* csr_frac;     // [nnz]
    float* out;                // [n_rows][T]
};

__global__ void combine_kernel(const __grid_constant__ CombineArgs a) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)a.n_rows * a.T) return;
    const int row = (int)(gid / a.T);
    const int t = (int)(gid % a.T);
    const int hop = a.n_hops > 0 ? row % a.n_hops : 0x7fffffff;
    float acc = 0.f, wsum = 0.f;
    for (int e = a.csr_ptr[t]; e < a.csr_ptr[t + 1]; ++e) {
        const int r = a.csr_res[e];
        const float* m = a.mag[r];
        if (m == nullptr || hop < a.first_frame[r]) continue;
        const float* mr = m + (size_t)row * a.bins[r];
        const int lo = a.csr_lo[e];
        const float m0 = mr[lo], m1 = mr[lo + 1];
        const float v = fmaf(m1 - m0, a.csr_frac[e], m0);
        acc += v * a.weight[r];
        wsum += a.weight[r];
    }
    a.out[gid] = wsum > 0.f ? acc / wsum : 0.f;
}

// ---------------------------------------------------------------------------------------------
// PrecomputedFrequencyMapper.map_spectrum_to_bars (freq_mapper.py:165-196)
// ---------------------------------------------------------------------------------------------
__global__ void band_map_kernel(const float* __restrict__ spec, int n_rows, int len,
                                const int* __restrict__ bands, int n_bars, int n_valid_bars,
                                const float* __restrict__ comp, float* __restrict__ out, int db) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_rows * n_bars) return;
    const int row = (int)(gid / n_bars);
    const int b = (int)(gid % n_bars);
    float val = 0.f;
    if (b < n_valid_bars) {
        const int s = bands[2 * b], e = bands[2 * b + 1];
        const float* sr = spec + (size_t)row * len;
        if (e > s) {
            float acc = 0.f;
            for (int i = s; i < e; ++i) acc += comp ? sr[i] * comp[i] : sr[i];
            val = acc / (float)(e - s);
        } else if (s < len) {
            val = comp ? sr[s] * comp[s] : sr[s];
        }
    }
    if (db == 1) val = 20.f * log10f(fmaxf(val, 1e-10f));           // panels/spectrogram_waterfall.py:85
    else if (db == 2) val = 20.f * log10f(val + 1e-10f);             // plugins/panels/spectrogram.py:72
    out[gid] = val;
}

// ---------------------------------------------------------------------------------------------
// Wire format of the capture side: interleaved little-endian int16 frames ->
// planar float32 rows,  x = int16 / 32768  (omega4/audio/capture.py:571-574; exact in float32).
// Stream g, frame i, channel c  ->  out[(g*il + c) * out_stride + i]
// ---------------------------------------------------------------------------------------------
__global__ void s16_deinterleave_kernel(const int16_t* __restrict__ in, long long stream_stride, int il,
                                        long long n_frames, float* __restrict__ out, long long out_stride) {
    const int g = blockIdx.y;
    const int16_t* src = in + (long long)g * stream_stride;
    const long long total = n_frames * il;
    const long long base = (long long)blockIdx.x * 1024;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long e = base + threadIdx.x + 256 * k;          // element index in the interleaved stream
        if (e < total) {
            const long long i = e / il;
            const int c = (int)(e - i * il);
            out[((long long)g * il + c) * out_stride + i] = (float)src[e] * (1.0f / 32768.0f);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// BassZoomPanel._process_bass_detail_internal (omega4/panels/bass_zoom.py:141-214), data side without the
// wall-clock peak hold:  raw = mean(|X|[bins of the bar]) * comp[bar];  scale = 0.85 / max(raw) *
// log10(max(1, 10 max(raw))) / 2;  v = raw * scale, compressed above 0.7 (0.7 + 0.3 (v - 0.7));  attack /
// release smoothing against the previous bars (0.1/0.9 when rising, 0.6/0.4 when falling);  clamp [0, 1].
// One warp per channel walks its frames in order (the smoothing is the only sequential dependency).
// ---------------------------------------------------------------------------------------------
__global__ void bass_bars_kernel(const float* __restrict__ mag, int n_ch, int n_frames, int n_bins,
                                 const int* __restrict__ bar_bins, const float* __restrict__ comp, int n_bars,
                                 float* __restrict__ state, float* __restrict__ out) {
    const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (ch >= n_ch) return;
    constexpr int MAXB = 4;                                    // up to 128 bars
    float prev[MAXB];
#pragma unroll
    for (int i = 0; i < MAXB; ++i) { const int b = lane + 32 * i; prev[i] = (state && b < n_bars) ? state[(size_t)ch * n_bars + b] : 0.f; }
    for (int f = 0; f < n_frames; ++f) {
        const float* row = mag + ((size_t)ch * n_frames + f) * n_bins;
        float raw[MAXB];
        float mx = 0.f;
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
            const int b = lane + 32 * i;
            raw[i] = -1.f;                                     // bars without bins keep their value (:163, :200)
            if (b < n_bars) {
                const int s = bar_bins[2 * b], n = bar_bins[2 * b + 1];
                if (n > 0) {
                    float acc = 0.f;
                    for (int k = 0; k < n; ++k) acc += row[s + k];
                    raw[i] = (acc / (float)n) * comp[b];
                    mx = fmaxf(mx, raw[i]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float scale = mx > 0.f ? (0.85f / mx) * (log10f(fmaxf(1.0f, mx * 10.f)) * 0.5f) : 1.0f;
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
            const int b = lane + 32 * i;
            if (b < n_bars) {
                float v = prev[i];
                if (raw[i] >= 0.f) {
                    const float sv = raw[i] * scale;
                    const float cv = sv > 0.7f ? 0.7f + (sv - 0.7f) * 0.3f : sv;
                    v = (cv > v) ? v * 0.1f + cv * 0.9f : v * 0.6f + cv * 0.4f;
                    v = fmaxf(0.f, fminf(1.f, v));
                }
                prev[i] = v;
                out[((size_t)ch * n_frames + f) * n_bars + b] = v;
            }
        }
    }
    if (state) {
#pragma unroll
        for (int i = 0; i < MAXB; ++i) { const int b = lane + 32 * i; if (b < n_bars) state[(size_t)ch * n_bars + b] = prev[i]; }
    }
}

// ---------------------------------------------------------------------------------------------
// Synthetic benchmark audio: log sweep 20 Hz -> 20 kHz over the clip, amplitude 0.5, start phase
// from the (stream, channel) hash of omega4_b200/batch/synth.py, plus white counter-hash noise of
// RMS 0.1 (the numpy generator shapes its noise pink; spectral colour is irrelevant to throughput).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}

__global__ void synth_kernel(float* __restrict__ out, int n_rows, int n_channels, long long n_samples,
                             long long row_stride, int first_stream, double sample_rate, double clip_seconds) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (i >= n_samples || row >= n_rows) return;
    const int stream = first_stream + row / n_channels;
    const int channel = row % n_channels;
    const uint32_t h = mix32((uint32_t)stream * 0x9E3779B1u + (uint32_t)channel * 0x85EBCA77u + 0x165667B1u);
    const double two_pi = 6.283185307179586476925287;
    const double k = log(20000.0 / 20.0) / clip_seconds;
    const double t = (double)i / sample_rate;
    const double phase = two_pi * 20.0 * expm1(k * t) / k + two_pi * (double)h / 4294967296.0;
    double sv = 0.5 * sin(phase);
    // four uniforms -> approximately gaussian, variance 4/12
    uint32_t c = mix32((uint32_t)(i & 0xffffffffu) ^ mix32(h + (uint32_t)(i >> 32) + 0x9E3779B9u));
    float u = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { c = mix32(c + 0x632BE5ABu); u += (float)(c >> 8) * (1.0f / 16777216.0f); }
    const float noise = (u - 2.0f) * (0.1f / 0.57735026919f);
    out[(long long)row * row_stride + i] = (float)sv + noise;
}

}  // namespace o4
