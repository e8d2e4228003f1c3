// multires_kernel.cuh -- batched windowed real FFT with fused |X| * weight epilogue and fused
// combine (np.interp onto the linear target grid) for ONE resolution.
//
// Replaces, for a whole batch of (channel, hop) frames at once:
//   MultiResolutionFFT.process_audio_chunk   (omega4/audio/multi_resolution_fft.py:251-286)
//   MultiResolutionFFT._apply_psychoacoustic_weighting (:304-333, weights precomputed on host)
//   MultiResolutionFFT.combine_results_optimized       (:359-395, tables precomputed on host)
//   BatchedFFTProcessor._process_size_group_cpu/gpu     (omega4/optimization/batched_fft_processor.py:197-285)
//   GPUAcceleratedFFT.compute_fft / process_fft_batch   (omega4/optimization/gpu_accelerated_fft.py:92-177,300-340)
//
// Two kernels share the epilogue:
//   multires_local_kernel<LOG2M>  N = 512 .. 8192   (fft_forward_local: 2 CTA barriers per round)
//   multires_kernel<LOG2M>        N = 16384, 32768  (generic multi-stage Stockham)
#pragma once
#include "fft_core.cuh"

namespace o4 {

struct MultiresArgs {
    const float* x;            // samples; frame f of channel c starts at x + c*ch_stride + f*frame_stride + frame_off0
    long long ch_stride;
    long long frame_stride;
    long long frame_off0;      // hop mode: hop - N (negative: reaches into history before x)
    int n_ch;
    int n_frames;              // frames per channel
    int first_frame;           // frames < first_frame are "not filled yet": zero outputs
    int rounds;                // rounds per CTA (each round = CONC frames)
    const float* window;       // [N] or nullptr
    const float* binw;         // [M+1] per-bin weights or nullptr
    const float2* twM;         // [M]      exp(-2 pi i e / M)
    const float2* twN;         // [M/2+1]  exp(-2 pi i k / N)
    float* mag_out;            // [n_ch][n_frames][M+1] or nullptr
    float2* cplx_out;          // [n_ch][n_frames][M+1] or nullptr
    float* comb_out;           // [n_ch][n_frames][T] or nullptr
    int T;
    int n_tb;                  // target bins written by this resolution
    const int* tb_idx;         // [n_tb] target bin index
    const int* tb_lo;          // [n_tb] lower FFT bin (or -1: write 0)
    const float* tb_frac;      // [n_tb] interpolation fraction
    float wnum, wden;          // out = interp * wnum / wden
    int need_lo, need_cnt;     // FFT bins [need_lo, need_lo + need_cnt) feed the combine step
};

struct ZPadded { static constexpr bool PADDED = true; __device__ static __forceinline__ int at(int i) { return padi(i); } };
template <int LOG2M>
struct ZSwizzled { static constexpr bool PADDED = false; __device__ static __forceinline__ int at(int i) { return zaddr<LOG2M>(i); } };

// Untangle + |X| * weight for the bins this thread owns: pairs u = t + i*TPF in [0, M/2), plus
// u = M/2 handled by thread 0.  In fused mode (no mag/complex output) only the bins the combine
// step needs are evaluated.
// twn / bw: twN and per-bin weight tables (global via the read-only path, or a shared-memory copy)
template <int LOG2M, typename ZA>
__device__ __forceinline__ void multires_epilogue(const MultiresArgs& a, const float2* Z, float* mags, int t,
                                                  size_t row, bool active, const float2* twn, const float* bw) {
    using S = FftShape<LOG2M>;
    constexpr int M = S::M, TPF = S::TPF;
    const bool all_bins = (a.mag_out != nullptr) || (a.cplx_out != nullptr);
    const int need_hi = a.need_lo + a.need_cnt;
    const int zt0 = ZA::at(t);
    const int zm0 = (t == 0) ? (ZA::PADDED ? M + M / 16 : M) : ZA::at(M - t);
    float* mrow = a.mag_out ? a.mag_out + row * (M + 1) : nullptr;
    float2* crow = a.cplx_out ? a.cplx_out + row * (M + 1) : nullptr;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const int u = (i < 8) ? t + i * TPF : M / 2;
        if (i == 8 && t != 0) break;
        const int k = u, km = M - u;
        const bool ink = (k >= a.need_lo && k < need_hi);
        const bool inm = (km >= a.need_lo && km < need_hi);
        if (!(all_bins || ink || inm)) continue;
        float mk = 0.f, mm = 0.f;
        float2 Xk = make_float2(0.f, 0.f), Xm = Xk;
        if (active) {
            const float2 Zk = Z[(i < 8) ? zt0 + i * TPF + (ZA::PADDED ? i * (TPF / 16) : 0) : ZA::at(u)];   // at(t + i TPF) is linear in i
            // mirrored bin: at(M - t - i TPF) = at(M - t) - i TPF (- i TPF / 16 when padded) for i >= 1; i = 0 wraps for t = 0
            const float2 Zm = Z[(i >= 1 && i < 8) ? zm0 - i * TPF - (ZA::PADDED ? i * (TPF / 16) : 0) : ZA::at((M - u) & (M - 1))];
            const float2 w = twn[u];
            rfft_pair(Zk, Zm, w, Xk, Xm);
            if (all_bins || ink) { mk = cabs_fast(Xk); if (bw) mk *= bw[k]; }
            if (all_bins || inm) { mm = cabs_fast(Xm); if (bw) mm *= bw[km]; }
        }
        if (mrow) { mrow[k] = mk; mrow[km] = mm; }
        if (crow) { crow[k] = Xk; crow[km] = Xm; }
        if (ink) mags[k - a.need_lo] = mk;
        if (inm) mags[km - a.need_lo] = mm;
    }
}

// Fused-mode epilogue (no magnitude / complex outputs): only the bins the combine step reads.  Whether iteration i
// touches a needed bin at all is decided per CTA from the uniform range (no divergent branches: the per-lane part is
// two predicated stores), and the mirrored half is skipped entirely when the needed range lies in the lower half of
// the spectrum (N = 2048: bins 43 .. 213 of 1025).
template <int LOG2M, typename ZA>
__device__ __forceinline__ void multires_epilogue_fused(const MultiresArgs& a, const float2* Z, float* mags, int t,
                                                        bool active, const float2* twn, const float* bw) {
    using S = FftShape<LOG2M>;
    constexpr int M = S::M, TPF = S::TPF;
    const int lo = a.need_lo, hi = a.need_lo + a.need_cnt;
    const unsigned cnt = (unsigned)a.need_cnt;
    const int zt0 = ZA::at(t);
    const int zm0 = (t == 0) ? (ZA::PADDED ? M + M / 16 : M) : ZA::at(M - t);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool anyk = ((i + 1) * TPF > lo) && (i * TPF < hi);                  // k  in [i TPF, (i+1) TPF)
        const bool anym = (M - i * TPF >= lo) && (M - (i + 1) * TPF + 1 < hi);      // km in [M - (i+1) TPF + 1, M - i TPF]
        if (!(anyk || anym)) continue;                                             // warp-uniform
        const int u = t + i * TPF, km = M - u;
        float mk = 0.f, mm = 0.f;
        if (active) {                                                              // warp-uniform
            const float2 Zk = Z[zt0 + i * TPF + (ZA::PADDED ? i * (TPF / 16) : 0)];
            const float2 Zm = Z[(i >= 1) ? zm0 - i * TPF - (ZA::PADDED ? i * (TPF / 16) : 0) : ZA::at((M - u) & (M - 1))];
            float2 Xk, Xm;
            rfft_pair(Zk, Zm, twn[u], Xk, Xm);
            if (anyk) { mk = cabs_fast(Xk); if (bw) mk *= bw[u]; }
            if (anym) { mm = cabs_fast(Xm); if (bw) mm *= bw[km]; }
        }
        if ((unsigned)(u - lo) < cnt) mags[u - lo] = mk;
        if ((unsigned)(km - lo) < cnt) mags[km - lo] = mm;
    }
    if (t == 0 && (unsigned)(M / 2 - lo) < cnt) {                                  // the self-paired bin k = M/2
        float mk = 0.f;
        if (active) {
            const float2 Zh = Z[ZA::at(M / 2)];
            float2 Xk, Xm;
            rfft_pair(Zh, Zh, twn[M / 2], Xk, Xm);
            mk = cabs_fast(Xk);
            if (bw) mk *= bw[M / 2];
        }
        mags[M / 2 - lo] = mk;
    }
}

// np.interp segments of combine_results_optimized evaluated from the smem strip of magnitudes.
struct CombTables { const int* idx; const int* lo; const float* frac; };

template <int TPF>
__device__ __forceinline__ void multires_combine(const MultiresArgs& a, const float* mags, int t, size_t row, bool active,
                                                 const CombTables tb) {
    float* orow = a.comb_out + row * a.T;
    // single-owner target bins: (interp * weight) / weight of the reference's accumulate / normalise is interp to 1 ulp
    const bool unit = (a.wnum == a.wden);
    for (int j = t; j < a.n_tb; j += TPF) {
        const int lo = tb.lo[j];
        float val = 0.f;
        if (active && lo >= 0) {
            const float m0 = mags[lo - a.need_lo];
            const float m1 = mags[lo + 1 - a.need_lo];
            const float vi = fmaf(m1 - m0, tb.frac[j], m0);
            val = unit ? vi : __fdividef(vi * a.wnum, a.wden);      // (interp * weight) / weight of :389-395
        }
        orow[tb.idx[j]] = val;
    }
}

// ---------------------------------------------------------------------------------------------
// N = 512 .. 8192
// ---------------------------------------------------------------------------------------------
template <int LOG2M>
__global__ void __launch_bounds__(256, 2)
multires_local_kernel(const __grid_constant__ MultiresArgs a) {
    using S = FftShape<LOG2M>;
    constexpr int M = S::M, TPF = S::TPF, CONC = S::CONC, BUF = S::BUF;
    static_assert(S::NT == 256, "local kernel runs 256-thread CTAs");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* bufs = reinterpret_cast<float2*>(smem_raw);                            // [CONC][BUF + M]
    float* mags_all = reinterpret_cast<float*>(bufs + (size_t)CONC * (BUF + M));   // [2][CONC][need_cnt]

    const int tid = threadIdx.x;
    const int g = tid / TPF;
    const int t = tid % TPF;
    float2* X = bufs + (size_t)g * (BUF + M);
    float2* Z = X + BUF;
    const int mstride = a.need_cnt > 0 ? a.need_cnt : 1;
    float* mags0 = mags_all + (size_t)g * mstride;
    float* mags1 = mags0 + (size_t)CONC * mstride;
    // small read-only tables are staged in shared memory once per CTA: ncu showed the N = 1024
    // kernel stalled 35 % of the time on their global (L1/L2) loads in the epilogue / combine
    int* s_tb_idx = reinterpret_cast<int*>(mags_all + (size_t)2 * CONC * mstride);
    int* s_tb_lo = s_tb_idx + a.n_tb;
    float* s_tb_frac = reinterpret_cast<float*>(s_tb_lo + a.n_tb);
    constexpr bool STAGE_TW = (LOG2M <= 10);
    float2* s_twn = reinterpret_cast<float2*>(s_tb_frac + a.n_tb + ((3 * a.n_tb) & 1));     // 8-byte aligned
    float* s_bw = reinterpret_cast<float*>(s_twn + (STAGE_TW ? M / 2 + 1 : 0));
    for (int j = threadIdx.x; j < a.n_tb; j += 256) {
        s_tb_idx[j] = __ldg(a.tb_idx + j); s_tb_lo[j] = __ldg(a.tb_lo + j); s_tb_frac[j] = __ldg(a.tb_frac + j);
    }
    if (STAGE_TW) {
        for (int j = threadIdx.x; j <= M / 2; j += 256) s_twn[j] = __ldg(a.twN + j);
        if (a.binw) for (int j = threadIdx.x; j <= M; j += 256) s_bw[j] = __ldg(a.binw + j);
    }
    __syncthreads();
    const CombTables ctb{s_tb_idx, s_tb_lo, s_tb_frac};
    const float2* twn = STAGE_TW ? s_twn : a.twN;
    const float* bw = a.binw ? (STAGE_TW ? s_bw : a.binw) : nullptr;

    const int frames_per_cta = a.rounds * CONC;
    const int tiles_per_ch = (a.n_frames + frames_per_cta - 1) / frames_per_cta;
    const int ch = blockIdx.x / tiles_per_ch;
    const int tile = blockIdx.x % tiles_per_ch;
    const int f0 = tile * frames_per_cta;

    // per-thread constants: all 2 x 15 stage twiddles in registers; the window pairs of the 16 owned
    // inputs are re-read from L1 every round (16 loads instead of 88 twiddle-product instructions)
    LocalTwFull<LOG2M> st;
    {
        LocalTw<LOG2M> st4;
        load_local_twiddles<LOG2M>(st4, a.twM, t);
        expand_local_twiddles<LOG2M>(st4, st);
    }
    const float2* wptr = reinterpret_cast<const float2*>(a.window) + t;
    // fused mode: the epilogue untangles only the bins k in [need_lo, need_lo + need_cnt) and needs Z[k] and Z[M - k] for
    // them; the 256-point blocks of the spectrum nobody reads are not stored (N = 2048: 2 of 4 blocks)
    unsigned zmask = 0xffffffffu;
    if (a.mag_out == nullptr && a.cplx_out == nullptr && a.need_cnt > 0) {
        zmask = 0u;
        const int lo = a.need_lo, hi = a.need_lo + a.need_cnt - 1;          // inclusive
        const int mlo = M - hi, mhi = M - lo;                               // mirrored partners (index M wraps to 0)
        for (int pq = 0; pq < (M / 256 > 0 ? M / 256 : 1); ++pq) {
            const int b0 = 256 * pq, b1 = b0 + 255;
            if ((lo <= b1 && hi >= b0) || (mlo <= b1 && mhi >= b0) || (pq == 0 && mhi >= M)) zmask |= 1u << pq;
        }
    }

    const float* xch = a.x + (long long)ch * a.ch_stride + a.frame_off0;
    float2 v[16];
    {
        const int f = f0 + g;
        if (f < a.n_frames && f >= a.first_frame) {
            const float2* px = reinterpret_cast<const float2*>(xch + (long long)f * a.frame_stride);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __ldg(px + t + j * TPF);
        }
    }

    for (int r = 0; r < a.rounds; ++r) {
        const int f = f0 + r * CONC + g;
        const bool valid = f < a.n_frames;
        const bool active = valid && f >= a.first_frame;
        if (active && a.window) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = cmul_elem(v[j], __ldg(wptr + j * TPF));
        }
        // the combine of the previous round runs right after this round's first barrier, which
        // also publishes the previous epilogue's magnitudes
        fft_forward_local<LOG2M, false>(v, X, Z, st, t, g, active, [&]() {
            if (r > 0 && a.comb_out) {
                const int fp = f - CONC;
                if (fp < a.n_frames)
                    multires_combine<TPF>(a, ((r - 1) & 1) ? mags1 : mags0, t, (size_t)ch * a.n_frames + fp,
                                          fp >= a.first_frame, ctb);
            }
        }, zmask);
        // prefetch the next round's samples; they land while the epilogue runs
        {
            const int fn = f + CONC;
            if (r + 1 < a.rounds && fn < a.n_frames && fn >= a.first_frame) {
                const float2* px = reinterpret_cast<const float2*>(xch + (long long)fn * a.frame_stride);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __ldg(px + t + j * TPF);
            }
        }
        if (valid) {
            if (a.mag_out == nullptr && a.cplx_out == nullptr)
                multires_epilogue_fused<LOG2M, ZSwizzled<LOG2M>>(a, Z, (r & 1) ? mags1 : mags0, t, active, twn, bw);
            else
                multires_epilogue<LOG2M, ZSwizzled<LOG2M>>(a, Z, (r & 1) ? mags1 : mags0, t, (size_t)ch * a.n_frames + f, active,
                                                           twn, bw);
        }
    }
    if (a.comb_out) {
        group_sync<TPF>(g);
        const int fl = f0 + (a.rounds - 1) * CONC + g;
        if (fl < a.n_frames)
            multires_combine<TPF>(a, ((a.rounds - 1) & 1) ? mags1 : mags0, t, (size_t)ch * a.n_frames + fl,
                                  fl >= a.first_frame, ctb);
    }
}

template <int LOG2M>
inline size_t multires_local_smem_bytes(int need_cnt, int n_tb) {
    using S = FftShape<LOG2M>;
    size_t b = (size_t)S::CONC * (S::BUF + S::M) * sizeof(float2) +
               (size_t)2 * S::CONC * (need_cnt > 0 ? need_cnt : 1) * sizeof(float) + (size_t)(3 * n_tb + 2) * 4;
    if (LOG2M <= 10) b += (size_t)(S::M / 2 + 1) * sizeof(float2) + (size_t)(S::M + 1) * sizeof(float);
    return b + 16;
}

// ---------------------------------------------------------------------------------------------
// N = 16384, 32768 (generic multi-stage path)
// ---------------------------------------------------------------------------------------------
template <int LOG2M>
__global__ void __launch_bounds__(FftShape<LOG2M>::NT, (FftShape<LOG2M>::NT <= 256 ? 2 : 1))
multires_kernel(const __grid_constant__ MultiresArgs a) {
    using S = FftShape<LOG2M>;
    constexpr int TPF = S::TPF, CONC = S::CONC, BUF = S::BUF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* bufs = reinterpret_cast<float2*>(smem_raw);
    float* mags_all = reinterpret_cast<float*>(bufs + (size_t)CONC * S::NBUF * BUF);

    const int tid = threadIdx.x;
    const int g = tid / TPF;
    const int t = tid % TPF;
    float2* buf0 = bufs + (size_t)g * S::NBUF * BUF;
    float2* buf1 = S::PINGPONG ? buf0 + BUF : buf0;
    float* mags = mags_all + (size_t)g * (a.need_cnt > 0 ? a.need_cnt : 1);

    const int frames_per_cta = a.rounds * CONC;
    const int tiles_per_ch = (a.n_frames + frames_per_cta - 1) / frames_per_cta;
    const int ch = blockIdx.x / tiles_per_ch;
    const int tile = blockIdx.x % tiles_per_ch;
    const int f0 = tile * frames_per_cta;

    float2 win[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
        win[j] = a.window ? __ldg(reinterpret_cast<const float2*>(a.window) + t + j * TPF) : make_float2(1.f, 1.f);
    StageTw<LOG2M> st;
    load_stage_twiddles<LOG2M>(st, a.twM, t);

    const float* xch = a.x + (long long)ch * a.ch_stride + a.frame_off0;
    float2 v[16];
    {
        const int f = f0 + g;
        if (f < a.n_frames && f >= a.first_frame) {
            const float2* px = reinterpret_cast<const float2*>(xch + (long long)f * a.frame_stride);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __ldg(px + t + j * TPF);
        }
    }

    for (int r = 0; r < a.rounds; ++r) {
        const int f = f0 + r * CONC + g;
        const bool valid = f < a.n_frames;
        const bool active = valid && f >= a.first_frame;
        if (active) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = cmul_elem(v[j], win[j]);
        }
        const float2* Z = fft_forward<LOG2M, false>(v, buf0, buf1, st, t, active);
        {
            const int fn = f + CONC;
            if (r + 1 < a.rounds && fn < a.n_frames && fn >= a.first_frame) {
                const float2* px = reinterpret_cast<const float2*>(xch + (long long)fn * a.frame_stride);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __ldg(px + t + j * TPF);
            }
        }
        const size_t row = (size_t)ch * a.n_frames + (valid ? f : 0);
        if (valid) multires_epilogue<LOG2M, ZPadded>(a, Z, mags, t, row, active, a.twN, a.binw);
        __syncthreads();
        if (valid && a.comb_out) multires_combine<TPF>(a, mags, t, row, active, CombTables{a.tb_idx, a.tb_lo, a.tb_frac});
        // No second barrier needed: every thread has finished reading Z before the barrier above,
        // and the next epilogue's writes to mags are separated from this combine's reads by the
        // barriers inside fft_forward.
    }
}

template <int LOG2M>
inline size_t multires_smem_bytes(int need_cnt) {
    using S = FftShape<LOG2M>;
    return (size_t)S::CONC * S::NBUF * S::BUF * sizeof(float2) + (size_t)S::CONC * (need_cnt > 0 ? need_cnt : 1) * sizeof(float);
}

}  // namespace o4
