// truepeak_kernel.cuh -- 4x oversampled true peak by EXACT periodic (FFT) interpolation.
//
// Replaces ProfessionalMetering.calculate_true_peak (omega4/panels/professional_meters.py:283-299),
// i.e. scipy.signal.resample(x, 4*len(x)) -> max|.| -> 20 log10, for a whole batch of meter frames.
// scipy's FFT-method resample of a real, even-length frame is: X = rfft(x); X[W/2] *= 0.5;
// irfft(zero-padded X, 4W) * 4.  The pruned form used here (max error
// 1.1e-15 against scipy in float64):  out[4n] = x[n], and for phase p in {1,2,3}
//     out[4n+p] = irfft_W(Xp)[n],  Xp[k] = X[k] e^{+2 pi i k p/(4W)} (k < W/2),  Xp[W/2] = X[W/2] cos(pi p/4)
// i.e. one forward and three inverse W-point real transforms, each done as a W/2-point complex
// Stockham transform (fft_core.cuh).  Only max|.| of the inverse outputs is needed, so the last
// stage of every inverse transform stays in registers.
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
    int rounds;
    const float* window;       // [W] float32 Hann (hop mode) or nullptr
    const float2* twM;         // [M]
    const float2* twN;         // [M/2+1]  exp(-2 pi i k / W)
    const float2* tw4W;        // [3M+1]   exp(+2 pi i j / (4W))
    double* tp_out;            // [n_ch][n_frames] dBTP
};

template <int LOG2M>
__global__ void __launch_bounds__(256, 2)
truepeak_kernel(const __grid_constant__ TruePeakArgs a) {
    using S = FftShape<LOG2M>;
    constexpr int M = S::M, TPF = S::TPF, CONC = S::CONC, BUF = S::BUF;
    constexpr int WARPS_PER_FFT = (TPF + 31) / 32;
    static_assert(TPF >= 32 && S::NT == 256 && LOG2M <= 12, "true-peak kernel: whole warps per sub-FFT, local FFT variant");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* bufs = reinterpret_cast<float2*>(smem_raw);                            // [CONC][BUF + M]
    float2* xs_all = bufs + (size_t)CONC * (BUF + M);                              // [CONC][M+1] spectrum
    float* red_all = reinterpret_cast<float*>(xs_all + (size_t)CONC * (M + 1));   // [CONC][WARPS_PER_FFT]

    const int tid = threadIdx.x;
    const int g = tid / TPF;
    const int t = tid % TPF;
    float2* X = bufs + (size_t)g * (BUF + M);
    float2* Z = X + BUF;
    float2* Xs = xs_all + (size_t)g * (M + 1);
    float* red = red_all + g * WARPS_PER_FFT;

    const int frames_per_cta = a.rounds * CONC;
    const int tiles_per_ch = (a.n_frames + frames_per_cta - 1) / frames_per_cta;
    const int ch = blockIdx.x / tiles_per_ch;
    const int tile = blockIdx.x % tiles_per_ch;
    const int f0 = tile * frames_per_cta;

    // all 2 x 15 stage twiddles live in registers (four transforms per frame reuse them); the window is
    // re-read from L1 once per frame instead
    LocalTwFull<LOG2M> st;
    {
        LocalTw<LOG2M> st4;
        load_local_twiddles<LOG2M>(st4, a.twM, t);
        expand_local_twiddles<LOG2M>(st4, st);
    }
    const float inv_m = 1.0f / (float)M;
    auto nop = []() {};

    for (int r = 0; r < a.rounds; ++r) {
        const int f = f0 + r * CONC + g;
        const bool active = (f < a.n_frames) && (f >= a.first_frame);
        float2 v[16];
        float pk = 0.f;                                   // phase 0: the frame itself
        if (active) {
            const long long off = (long long)ch * a.ch_stride + a.frame_off0 + (long long)f * a.frame_stride;
            if (a.x_is_f64) {
                const double2* px = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(a.x) + off);
#pragma unroll
                for (int j = 0; j < 16; ++j) { double2 d = px[t + j * TPF]; v[j] = make_float2((float)d.x, (float)d.y); }
            } else {
                const float2* px = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(a.x) + off);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __ldg(px + t + j * TPF);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 w = a.window ? __ldg(reinterpret_cast<const float2*>(a.window) + t + j * TPF) : make_float2(1.f, 1.f);
                v[j].x *= w.x; v[j].y *= w.y;
                pk = fmaxf(pk, fmaxf(fabsf(v[j].x), fabsf(v[j].y)));
            }
        }
        fft_forward_local<LOG2M, false>(v, X, Z, st, t, g, active, nop);
        // untangle -> X[0..M] in Xs
        if (active) {
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                int u = (i < 8) ? t + i * TPF : M / 2;
                if (i == 8 && t != 0) break;
                float2 Zk = Z[zaddr<LOG2M>(u)];
                float2 Zm = Z[zaddr<LOG2M>((M - u) & (M - 1))];
                float2 Xk, Xm;
                rfft_pair(Zk, Zm, __ldg(a.twN + u), Xk, Xm);
                Xs[u] = Xk;
                Xs[M - u] = Xm;
            }
        }
        group_sync<TPF>(g);
        float pko = 0.f;                                  // phases 1..3, unnormalised (x M)
#pragma unroll 1
        for (int p = 1; p <= 3; ++p) {
            if (active) {
                // e^{i pi p/4}
                const float rr = 0.70710678118654752440f;
                const float2 cp = (p == 1) ? make_float2(rr, rr) : (p == 2) ? make_float2(0.f, 1.f) : make_float2(-rr, rr);
                const float cosp = (p == 2) ? 0.f : (p == 1 ? rr : -rr);
                // Build c[k] = conj(Z'[k]) for the inverse transform, pair (k, M-k) at a time.
                //   A = Xp[k], B = Xp[M-k];  E = (A + conj B)/2;  O = (A - conj B)/2 * e^{+2 pi i k/W}
                //   Z'[k] = E + iO, Z'[M-k] = conj(E) + i conj(O)  ->  c[k] = conj(E) - i conj(O), c[M-k] = E - iO
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    int u = (i < 8) ? t + i * TPF : M / 2;
                    if (i == 8 && t != 0) break;
                    float2 A = Xs[u];
                    float2 B = Xs[M - u];
                    float2 P = __ldg(a.tw4W + u * p);                 // e^{+2 pi i u p/(4W)}
                    A = cmul(A, P);
                    if (u == 0) B = make_float2(B.x * cosp, 0.f);     // Nyquist bin of the W-point spectrum
                    else B = cmul(B, cmul(cp, make_float2(P.x, -P.y)));
                    float2 tw = __ldg(a.twN + u);                     // e^{-2 pi i u/W}; conj -> e^{+...}
                    float2 E = make_float2(0.5f * (A.x + B.x), 0.5f * (A.y - B.y));
                    float2 H = make_float2(0.5f * (A.x - B.x), 0.5f * (A.y + B.y));
                    float2 O = cmul(H, make_float2(tw.x, -tw.y));
                    // c[k] = conj(E) - i conj(O) = (E.x - O.y, -E.y - O.x);  c[M-k] = E - iO = (E.x + O.y, E.y - O.x)
                    Z[zaddr<LOG2M>(u)] = make_float2(E.x - O.y, -E.y - O.x);
                    if (u != 0) Z[zaddr<LOG2M>(M - u)] = make_float2(E.x + O.y, E.y - O.x);
                }
            }
            group_sync<TPF>(g);      // c[] complete; also: every thread is past the previous transform's X reads
            if (active) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = Z[zaddr<LOG2M>(t + j * TPF)];
            }
            fft_forward_local<LOG2M, true>(v, X, Z, st, t, g, active, nop);
            if (active) {
#pragma unroll
                for (int j = 0; j < 16; ++j) pko = fmaxf(pko, fmaxf(fabsf(v[j].x), fabsf(v[j].y)));
            }
        }
        pk = fmaxf(pk, pko * inv_m);
        // reduce over the sub-FFT's threads
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
        if ((t & 31) == 0) red[t >> 5] = pk;
        group_sync<TPF>(g);          // also orders the last transform's X reads before the next round's stage-1 store
        if (active && t == 0) {
            float m = red[0];
#pragma unroll
            for (int w = 1; w < WARPS_PER_FFT; ++w) m = fmaxf(m, red[w]);
            double peak = (double)m;
            a.tp_out[(size_t)ch * a.n_frames + f] = (peak < 1e-10) ? -100.0 : 20.0 * log10(peak);
        }
        // red[] is next written after several barriers of the next round
    }
}

template <int LOG2M>
inline size_t truepeak_smem_bytes() {
    using S = FftShape<LOG2M>;
    return (size_t)S::CONC * (S::BUF + S::M) * sizeof(float2) + (size_t)S::CONC * (S::M + 1) * sizeof(float2)
           + (size_t)S::CONC * ((S::TPF + 31) / 32) * sizeof(float) + 16;
}

}  // namespace o4
