// truepeak_kernel.cuh -- 4x oversampled true peak by EXACT periodic (FFT) interpolation.
//
// Replaces ProfessionalMetering.calculate_true_peak (omega4/panels/professional_meters.py:283-299),
// i.e. scipy.signal.resample(x, 4*len(x)) -> max|.| -> 20 log10, for a whole batch of meter frames.
// scipy's FFT-method resample of a real, even-length frame is: X = rfft(x); X[W/2] *= 0.5;
// irfft(zero-padded X, 4W) * 4.  The pruned form used here (max error
// 1.1e-15 against scipy in float64):  out[4n] = x[n], and for phase p in {1,2,3}
//     out[4n+p] = ifft_W(S_p)[n],   S_p[k] = X[k] e^{+2 pi i f(k) p/(4W)},  f(k) = k (k < W/2), k - W (k > W/2),
//     S_p[W/2] = X[W/2] cos(pi p/4)
// i.e. a quarter-, half- and three-quarter-sample delay applied in the frequency domain.
//
// TWO FRAMES PER TRANSFORM: consecutive frames a, b of a channel are packed as z = a + i b.  The delay
// factors depend on the signed frequency only, so S_p = Z .* R_p holds for the packed spectrum as it
// stands -- no separation of the two spectra, no real/complex (un)tangling: one forward and three
// "inverse" W-point complex transforms per PAIR (inverse = forward transform of the conjugate; only
// max|Re|, max|Im| of the outputs are needed, so the conjugation of the result is irrelevant and the
// last stage of every inverse transform stays in registers).  Each frame is first scaled to unit
// sample peak, so a quiet frame is not polluted by the fp32 rounding noise of a loud neighbour, and
// an all-zero frame returns -100 exactly as the reference does.
//
// A polyphase FIR cannot meet the 0.05 dBTP parity bar against this reference.
#pragma once
#include "fft_core.cuh"

namespace o4 {

struct TruePeakArgs {
    const void* x;             // float samples (hop mode) or double frames (frames mode)
    int x_is_f64;
    long long ch_stride;
    long long frame_stride;
    long long frame_off0;
    int n_ch;
    int n_frames;
    int first_frame;           // frames before this are not measured (tp_out untouched)
    int rounds;                // frame pairs per 128-thread group
    const float* window;       // [W] float32 Hann (hop mode) or nullptr
    const float2* twM;         // [W]      exp(-2 pi i e / W)
    const float2* rot;         // [3][W]   R_p[k], p = 1..3 (see above)
    const unsigned* rot_h;     // [3][W]   the same as packed half (re, im); truepeak16h_kernel only
    double* tp_out;            // [n_ch][n_frames] dBTP
    // R_p[t + j W/16] = R_p[t] * step[p-1][j]:  step = e^{+2 pi i p j'/64}, j' = j (j < 8) or j - 16 (signed frequency);
    // uniform constant-bank operands instead of 48 table loads per thread and pair
    float2 step[3][16];
    float nyq[3];              // cos(pi p / 4): factor of the bin k = W/2
};

template <int LOG2M>
__global__ void __launch_bounds__(256, 2)
truepeak_kernel(const __grid_constant__ TruePeakArgs a) {
    using S = FftShape<LOG2M>;
    constexpr int M = S::M, TPF = S::TPF, CONC = S::CONC, BUF = S::BUF;
    constexpr int WARPS = TPF / 32;
    static_assert(TPF >= 32 && S::NT == 256 && LOG2M <= 12, "true-peak kernel: whole warps per transform, local FFT variant");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* bufs = reinterpret_cast<float2*>(smem_raw);                            // [CONC][BUF + M]
    float2* red_all = bufs + (size_t)CONC * (BUF + M);                             // [CONC][2][WARPS] (a, b) maxima

    const int tid = threadIdx.x;
    const int g = tid / TPF;
    const int t = tid % TPF;
    float2* X = bufs + (size_t)g * (BUF + M);
    float2* Z = X + BUF;
    float2* red = red_all + g * 2 * WARPS;

    const int pairs_per_cta = a.rounds * CONC;
    const int n_pairs = (a.n_frames + 1) >> 1;
    const int tiles_per_ch = (n_pairs + pairs_per_cta - 1) / pairs_per_cta;
    const int ch = blockIdx.x / tiles_per_ch;
    const int tile = blockIdx.x % tiles_per_ch;

    LocalTwFull<LOG2M> st;
    {
        LocalTw<LOG2M> st4;
        load_local_twiddles<LOG2M>(st4, a.twM, t);
        expand_local_twiddles<LOG2M>(st4, st);
    }
    const float inv_m = 1.0f / (float)M;
    auto nop = []() {};
    // group-wide maximum of a float2 (component-wise); every thread of the group gets the result
    auto group_max2 = [&](float2 v, int slot) -> float2 {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v.x = fmaxf(v.x, __shfl_xor_sync(0xffffffffu, v.x, o));
            v.y = fmaxf(v.y, __shfl_xor_sync(0xffffffffu, v.y, o));
        }
        if ((t & 31) == 0) red[slot * WARPS + (t >> 5)] = v;
        group_sync<TPF>(g);
        float2 m = red[slot * WARPS];
#pragma unroll
        for (int w = 1; w < WARPS; ++w) { const float2 o = red[slot * WARPS + w]; m.x = fmaxf(m.x, o.x); m.y = fmaxf(m.y, o.y); }
        return m;
    };

    const int zt = zaddr<LOG2M>(t);
    // this thread's base delay factors R_p[t]
    const float2 rbase1 = __ldg(a.rot + t), rbase2 = __ldg(a.rot + M + t), rbase3 = __ldg(a.rot + 2 * M + t);
    // raw samples of a pair: v[j] = (a[t + j TPF], b[t + j TPF]); fetched one round ahead
    auto load_pair = [&](int pi, float2 (&v)[16]) {
        const int fa = 2 * pi, fb = fa + 1;
        const bool act_a = (fa < a.n_frames) && (fa >= a.first_frame);
        const bool act_b = (fb < a.n_frames) && (fb >= a.first_frame);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = make_float2(0.f, 0.f);
        if (!(act_a || act_b)) return;
        // one 64-bit base per frame; the 16 loads of a frame are then immediate offsets j * TPF from it
        const long long off_a = (long long)ch * a.ch_stride + a.frame_off0 + (long long)fa * a.frame_stride + t;
        if (a.x_is_f64) {
            const double* pa = reinterpret_cast<const double*>(a.x) + off_a;
            const double* pb = pa + a.frame_stride;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (act_a) v[j].x = (float)pa[j * TPF];
                if (act_b) v[j].y = (float)pb[j * TPF];
            }
        } else {
            const float* pa = reinterpret_cast<const float*>(a.x) + off_a;
            const float* pb = pa + a.frame_stride;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (act_a) v[j].x = __ldg(pa + j * TPF);
                if (act_b) v[j].y = __ldg(pb + j * TPF);
            }
        }
    };
    float2 v[16];
    load_pair(tile * pairs_per_cta + g, v);
    for (int r = 0; r < a.rounds; ++r) {
        const int pi = tile * pairs_per_cta + r * CONC + g;
        const int fa = 2 * pi, fb = fa + 1;
        const bool act_a = (fa < a.n_frames) && (fa >= a.first_frame);
        const bool act_b = (fb < a.n_frames) && (fb >= a.first_frame);
        const bool active = act_a || act_b;
        if (a.window) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { const float w = __ldg(a.window + t + j * TPF); v[j].x *= w; v[j].y *= w; }
        }
        // phase 0 = the frames themselves; scale each to unit sample peak
        float2 pk = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 16; ++j) { pk.x = fmaxf(pk.x, fabsf(v[j].x)); pk.y = fmaxf(pk.y, fabsf(v[j].y)); }
        pk = group_max2(pk, 0);
        // (a sample peak below 1e-20 can only end below the reference's 1e-10 floor: skip the scaling, whose
        // reciprocal would overflow for denormal peaks)
        const float sa = pk.x > 1e-20f ? 1.f / pk.x : 0.f, sb = pk.y > 1e-20f ? 1.f / pk.y : 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { v[j].x *= sa; v[j].y *= sb; }
        fft_forward_local<LOG2M, false>(v, X, Z, st, t, g, active, nop);          // ends with a group barrier
        float2 pko = make_float2(0.f, 0.f);                       // phases 1..3, scaled by M
#pragma unroll 1
        for (int p = 1; p <= 3; ++p) {
            if (active) {
                const float2 rb = (p == 1) ? rbase1 : (p == 2 ? rbase2 : rbase3);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int k = t + j * TPF;
                    float2 rk = cmul(rb, a.step[p - 1][j]);
                    if (j == 8 && t == 0) rk = make_float2(a.nyq[p - 1], 0.f);      // k = W/2
                    const float2 s = cmul(Z[zt + j * TPF], rk);       // zaddr(t + j TPF) = zaddr(t) + j TPF (TPF = 16 G2)
                    v[j] = make_float2(s.x, -s.y);                // conj: inverse transform by the forward kernel
                }
            }
            fft_forward_local<LOG2M, true>(v, X, Z, st, t, g, active, nop);
            if (active) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { pko.x = fmaxf(pko.x, fabsf(v[j].x)); pko.y = fmaxf(pko.y, fabsf(v[j].y)); }
            }
            group_sync<TPF>(g);      // every thread is past this transform's X reads before the next stage-1 store
        }
        // the next round's samples travel while the maxima are reduced and written
        if (r + 1 < a.rounds) load_pair(tile * pairs_per_cta + (r + 1) * CONC + g, v);
        pko = group_max2(pko, 1);
        if (t == 0) {
            const float na = fmaxf(1.f, pko.x * inv_m), nb = fmaxf(1.f, pko.y * inv_m);   // normalised peaks (phase 0 = 1)
            if (act_a) {
                const double peak = (double)pk.x * (double)na;
                a.tp_out[(size_t)ch * a.n_frames + fa] = (peak < 1e-10) ? -100.0 : 20.0 * log10(peak);
            }
            if (act_b) {
                const double peak = (double)pk.y * (double)nb;
                a.tp_out[(size_t)ch * a.n_frames + fb] = (peak < 1e-10) ? -100.0 : 20.0 * log10(peak);
            }
        }
        group_sync<TPF>(g);          // red[] and Z are reused by the next round
    }
}

template <int LOG2M>
inline size_t truepeak_smem_bytes() {
    using S = FftShape<LOG2M>;
    return (size_t)S::CONC * (S::BUF + S::M) * sizeof(float2) + (size_t)S::CONC * 2 * (S::TPF / 32) * sizeof(float2) + 16;
}

}  // namespace o4
