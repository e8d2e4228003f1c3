// blockdft_kernel.cuh -- "hop-block partial DFT": the fused path of the LONG resolutions when the
// combine step consumes only a handful of their FFT bins.
//
// combine_results_optimized (omega4/audio/multi_resolution_fft.py:353-391) interpolates the
// N = 8192 spectrum (20-200 Hz) onto 5 target bins: only 10 of its 4097 FFT bins are ever read
// (BASELINE config 5: 2 bins of the 32768-point and 8 of the 16384-point transform).  For those
// resolutions the windowed transform of process_audio_chunk (:251-286) is evaluated directly at
// the needed bins, with the work that all overlapping frames share done ONCE per hop block:
//
//   frame f, local index i = H b + n (block b < B = N/H, n < H = hop); block j = f + 1 - B + b
//   window  w[i] = sum_m a_m cos(m phi i), phi = 2 pi/(N-1)       (np.blackman / hamming / ones)
//   w[i] e^{-2 pi i k i/N} = sum_t c_t e^{i theta_t i},  theta_t = m_t phi - 2 pi k/N,  m_t in {0,+-1,+-2}
//   X_f[k] = sum_b sum_t  T[k][b][t] * Q_j[k][t]
//       Q_j[k][t]  = sum_n x_j[n] e^{i theta_t n}        <- blockdft_gemm_kernel: [blocks x H] x [H x cols]
//       T[k][b][t] = c_t e^{i theta_t H b}               <- blockdft_assemble_kernel (+ |X| * weight + np.interp)
//
// Cost per channel-hop for BASELINE's 8192: 100 real columns x 512 FMAs = 51 k FMA, against
// 266 k instructions of the full real FFT.  All arithmetic is fp32 on the CUDA cores (fp32 FFMA
// GEMM, 128 x BN tile, 8 x TN register tile, 3-stage cp.async pipeline); errors measured against
// numpy's float32 rfft: <= 1e-5 relative on bins 60 dB below the row maximum.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fft_core.cuh"

namespace o4 {

constexpr int BD_BM = 128;          // blocks (GEMM rows) per CTA
constexpr int BD_KC = 32;           // K chunk (samples) per pipeline stage
constexpr int BD_APAD = 4;          // A row padding (floats): row stride 36 words, conflict-free LDS.128
constexpr int BD_STAGES = 3;
constexpr int BD_THREADS = 256;     // 16 row groups x 16 column groups

struct BlockDftGemmArgs {
    const float* x;            // samples; block j of channel c = x + c*ch_stride + j*hop
    long long ch_stride;
    int hop;                   // K of the GEMM, multiple of BD_KC
    int n_ch;
    int j0;                    // first block (may be negative: history before x)
    int nb;                    // blocks per channel
    const float* E;            // [hop][BN] (cos, sin) columns, zero padded
    float* Q;                  // [n_ch][nb][BN]
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int BN>
__global__ void __launch_bounds__(BD_THREADS, 2)
blockdft_gemm_kernel(const __grid_constant__ BlockDftGemmArgs a) {
    constexpr int TN = BN / 16;                     // columns per thread: 8, 4 or 2
    constexpr int TM = 8;
    constexpr int ASTR = BD_KC + BD_APAD;
    static_assert(BN == 128 || BN == 64 || BN == 32, "column tile");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);                       // [STAGES][BM][ASTR]
    float* Bs = As + BD_STAGES * BD_BM * ASTR;                            // [STAGES][KC][BN]

    const int tid = threadIdx.x;
    const int rg = tid >> 4, cg = tid & 15;
    const int tiles_per_ch = (a.nb + BD_BM - 1) / BD_BM;
    const int ch = blockIdx.x / tiles_per_ch;
    const int row0 = (blockIdx.x % tiles_per_ch) * BD_BM;
    const float* xa = a.x + (long long)ch * a.ch_stride + (long long)a.j0 * a.hop;

    // cp.async assignments: A = 128 rows x 8 x 16 B, B = 32 rows x BN/4 x 16 B
    const int a_c4 = tid & 7, a_r = tid >> 3;       // rows a_r + 32 q
    auto load_stage = [&](int st, int kc) {
        float* as = As + st * BD_BM * ASTR;
        float* bs = Bs + st * BD_KC * BN;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int r = row0 + a_r + 32 * q;
            if (r >= a.nb) r = a.nb - 1;             // clamp: rows past the end are computed but not stored
            cp_async16(as + (a_r + 32 * q) * ASTR + a_c4 * 4, xa + (long long)r * a.hop + a_c4 * 4 + kc * BD_KC);
        }
        constexpr int B_VEC = BD_KC * BN / 4;        // 16-byte pieces per stage
#pragma unroll
        for (int i = 0; i < (B_VEC + BD_THREADS - 1) / BD_THREADS; ++i) {
            const int v = tid + i * BD_THREADS;
            if (B_VEC % BD_THREADS == 0 || v < B_VEC) {
                const int kr = v / (BN / 4), c4 = v % (BN / 4);
                cp_async16(bs + kr * BN + c4 * 4, a.E + (size_t)(kc * BD_KC + kr) * BN + c4 * 4);
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int n_kc = a.hop / BD_KC;
#pragma unroll
    for (int s = 0; s < BD_STAGES - 1; ++s) {
        if (s < n_kc) load_stage(s, s);
        cp_async_commit();
    }
    for (int kc = 0; kc < n_kc; ++kc) {
        cp_async_wait<BD_STAGES - 2>();
        __syncthreads();                                   // stage kc landed; stage kc-1 fully consumed
        {
            const int nx = kc + BD_STAGES - 1;
            if (nx < n_kc) load_stage(nx % BD_STAGES, nx);
            cp_async_commit();
        }
        const float* as = As + (kc % BD_STAGES) * BD_BM * ASTR + rg * ASTR;
        const float* bs = Bs + (kc % BD_STAGES) * BD_KC * BN;
#pragma unroll
        for (int k2 = 0; k2 < BD_KC / 2; ++k2) {
            float2 av[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) av[i] = *reinterpret_cast<const float2*>(as + 16 * i * ASTR + k2 * 2);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                float bv[TN];
                const float* brow = bs + (k2 * 2 + kk) * BN;
                if constexpr (TN == 8) {
                    const float4 b0 = *reinterpret_cast<const float4*>(brow + cg * 4);
                    const float4 b1 = *reinterpret_cast<const float4*>(brow + 64 + cg * 4);
                    bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
                    bv[4 % TN] = b1.x; bv[5 % TN] = b1.y; bv[6 % TN] = b1.z; bv[7 % TN] = b1.w;
                } else if constexpr (TN == 4) {
                    const float4 b0 = *reinterpret_cast<const float4*>(brow + cg * 4);
                    bv[0] = b0.x; bv[1] = b0.y; bv[2 % TN] = b0.z; bv[3 % TN] = b0.w;
                } else {
                    const float2 b0 = *reinterpret_cast<const float2*>(brow + cg * 2);
                    bv[0] = b0.x; bv[1] = b0.y;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const float av_k = (kk == 0) ? av[i].x : av[i].y;
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av_k, bv[j], acc[i][j]);
                }
            }
        }
    }
    cp_async_wait<0>();

    float* q = a.Q + ((size_t)ch * a.nb) * BN;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = row0 + rg + 16 * i;
        if (r >= a.nb) continue;
        float* qr = q + (size_t)r * BN;
        if constexpr (TN == 8) {
            *reinterpret_cast<float4*>(qr + cg * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            *reinterpret_cast<float4*>(qr + 64 + cg * 4) = make_float4(acc[i][4 % TN], acc[i][5 % TN], acc[i][6 % TN], acc[i][7 % TN]);
        } else if constexpr (TN == 4) {
            *reinterpret_cast<float4*>(qr + cg * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2 % TN], acc[i][3 % TN]);
        } else {
            *reinterpret_cast<float2*>(qr + cg * 2) = make_float2(acc[i][0], acc[i][1]);
        }
    }
}

template <int BN>
inline size_t blockdft_gemm_smem_bytes() {
    return (size_t)BD_STAGES * (BD_BM * (BD_KC + BD_APAD) + BD_KC * BN) * sizeof(float);
}

// ---------------------------------------------------------------------------------------------
// Assembly: X_f[k] = sum_b sum_t T[k][b][t] Q_{f+1-B+b}[k][t], then |X| * weight and the np.interp
// segments of combine_results_optimized straight into the combined row.
// One CTA = BD_FA consecutive frames of one channel; the Q rows they share are staged in smem.
// ---------------------------------------------------------------------------------------------
constexpr int BD_FA_MAX = 64;        // frames per CTA: 64, 32 or 16, whichever keeps the staged Q rows under ~40 KB

struct BlockDftAsmArgs {
    const float* Q;            // [n_ch][nb][qs]
    int qs;                    // Q row stride (floats)
    int col0;                  // first Q column of this resolution (even)
    int nb, j0;
    int nk, nt, B;             // needed bins, window terms, blocks per frame
    const float2* T;           // [nk][B][nt]
    const float* kw;           // [nk] per-bin weight (1 when weighting is off)
    int n_ch, n_frames, first_frame;
    float* comb_out;           // [n_ch][n_frames][Tbins]
    int Tbins;
    int n_tb;
    const int* tb_idx;         // [n_tb] target bin
    const int* tb_pos;         // [n_tb] position of the lower FFT bin in the needed-bin list, -1: write 0
    const float* tb_frac;      // [n_tb]
    float wnum, wden;
    int fa;                    // frames per CTA
};

// Register blocking: one work item = (needed bin ki, group of 4 consecutive frames).  The four frames
// share all but one Q row per block step, so a sliding window of four rows lives in registers and
// every step loads NT window terms + NT new Q values for 4 NT complex MACs (0.5 LDS per MAC instead
// of 2).  Lanes run over ki (Q stride NT float2: conflict free; T is stored [b][t][ki]).
template <int NT>
__global__ void __launch_bounds__(256)
blockdft_assemble_kernel(const __grid_constant__ BlockDftAsmArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nc = a.nk * NT;                             // complex columns of this resolution
    const int rs = nc | 1;                                // odd float2 row stride
    float2* Qs = reinterpret_cast<float2*>(smem_raw);     // [fa + B - 1 + 3][rs]
    float2* Ts = Qs + (size_t)(a.fa + a.B + 2) * rs;      // [B][NT][nk]
    float* mags = reinterpret_cast<float*>(Ts + (size_t)a.nk * a.B * NT);   // [fa][nk]

    const int tiles_per_ch = (a.n_frames + a.fa - 1) / a.fa;
    const int ch = blockIdx.x / tiles_per_ch;
    const int f0 = (blockIdx.x % tiles_per_ch) * a.fa;
    const int nf = min(a.fa, a.n_frames - f0);
    const int tid = threadIdx.x;

    for (int i = tid; i < a.nk * a.B * NT; i += 256) {    // global [ki][b][t] -> shared [b][t][ki]
        const int ki = i / (a.B * NT), r = i % (a.B * NT);
        Ts[r * a.nk + ki] = a.T[i];
    }
    // Q rows of blocks f0 + 1 - B .. (row index in Q = block - j0); rows before j0 belong to frames that
    // are not filled yet, rows past the channel's last block to the unused tail of the last frame group
    const int jb = f0 + 1 - a.B;
    const int nrows = a.fa + a.B + 2;
    const float* qch = a.Q + (size_t)ch * a.nb * a.qs + a.col0;
    for (int i = tid; i < nrows * nc; i += 256) {
        const int r = i / nc, c = i % nc;
        const int qr = jb + r - a.j0;
        float2 v = make_float2(0.f, 0.f);
        if (qr >= 0 && qr < a.nb) v = *reinterpret_cast<const float2*>(qch + (size_t)qr * a.qs + 2 * c);
        Qs[r * rs + c] = v;
    }
    __syncthreads();
    const int ng = (nf + 3) >> 2;
    for (int it = tid; it < ng * a.nk; it += 256) {
        const int ki = it % a.nk, fl = (it / a.nk) << 2;
        float2 acc[4], w[4][NT];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = make_float2(0.f, 0.f);
        const float2* qcol = Qs + fl * rs + ki * NT;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int t = 0; t < NT; ++t) w[j + 1][t] = qcol[j * rs + t];
        for (int b = 0; b < a.B; ++b) {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
#pragma unroll
                for (int j = 0; j < 3; ++j) w[j][t] = w[j + 1][t];
                w[3][t] = qcol[(b + 3) * rs + t];
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float2 tw = Ts[(b * NT + t) * a.nk + ki];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[j].x = fmaf(w[j][t].x, tw.x, fmaf(-w[j][t].y, tw.y, acc[j].x));
                    acc[j].y = fmaf(w[j][t].x, tw.y, fmaf(w[j][t].y, tw.x, acc[j].y));
                }
            }
        }
        const float kw = a.kw[ki];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (fl + j < nf) mags[(fl + j) * a.nk + ki] = sqrtf(fmaf(acc[j].x, acc[j].x, acc[j].y * acc[j].y)) * kw;
    }
    __syncthreads();
    for (int it = tid; it < nf * a.n_tb; it += 256) {
        const int fl = it / a.n_tb, j = it % a.n_tb;
        const int f = f0 + fl;
        const int pos = a.tb_pos[j];
        float val = 0.f;
        if (pos >= 0 && f >= a.first_frame) {
            const float m0 = mags[fl * a.nk + pos], m1 = mags[fl * a.nk + pos + 1];
            const float vi = fmaf(m1 - m0, a.tb_frac[j], m0);
            val = (vi * a.wnum) / a.wden;
        }
        a.comb_out[((size_t)ch * a.n_frames + f) * a.Tbins + a.tb_idx[j]] = val;
    }
}

inline size_t blockdft_assemble_smem_bytes(int nk, int nt, int B, int fa) {
    const int nc = nk * nt, rs = nc | 1;
    return (size_t)(fa + B + 2) * rs * sizeof(float2) + (size_t)nk * B * nt * sizeof(float2) +
           (size_t)fa * nk * sizeof(float) + 16;
}

inline int blockdft_assemble_frames(int nk, int nt, int B) {
    int fa = BD_FA_MAX;
    while (fa > 16 && blockdft_assemble_smem_bytes(nk, nt, B, fa) > 40 * 1024) fa >>= 1;
    return fa;
}

// ---------------------------------------------------------------------------------------------
// Time-domain-windowed variant ("exact windowing"): the GEMM operand already carries the window,
//   Q_j[k][b] = sum_n x_j[n] w[H b + n] e^{-2 pi i k (H b + n) / N}        one column pair per (bin, block position b)
// so a frame is a plain sum over the block positions, X_f[k] = sum_b Q_{f+1-B+b}[k][b], with no
// cancellation between terms: the rounding error scales with the WINDOWED energy of the frame exactly as
// in a time-domain windowed FFT (the cosine-sum formulation above multiplies the unwindowed block spectra
// by 0.42 / -0.25 / 0.04 and loses up to 0.03 dB on a burst that sits under the window's near-zero edge).
// It costs B/nt times the GEMM columns (BASELINE: 320 + 640 instead of 100 + 400), works for ANY window,
// and its assembly is this kernel: stage the Q rows of fa + B - 1 blocks, lanes over frames (odd float2 row
// stride: conflict free), then |X| * weight and the np.interp segments as above.
// ---------------------------------------------------------------------------------------------
struct BlockDftSumArgs {
    const float* Q;            // [n_ch][nb][qs]
    int qs, col0, nb, j0;
    int nk, B;
    const float* kw;           // [nk]
    int n_ch, n_frames, first_frame;
    float* comb_out;
    int Tbins, n_tb;
    const int* tb_idx;
    const int* tb_pos;
    const float* tb_frac;
    float wnum, wden;
    int fa;
};

__global__ void __launch_bounds__(256)
blockdft_sum_kernel(const __grid_constant__ BlockDftSumArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nc = a.nk * a.B;                            // complex columns of this resolution
    const int rs = nc | 1;                                // odd float2 row stride
    const int nrows = a.fa + a.B - 1;
    float2* Qs = reinterpret_cast<float2*>(smem_raw);     // [nrows][rs]
    float* mags = reinterpret_cast<float*>(Qs + (size_t)nrows * rs);   // [fa][nk]

    const int tiles_per_ch = (a.n_frames + a.fa - 1) / a.fa;
    const int ch = blockIdx.x / tiles_per_ch;
    const int f0 = (blockIdx.x % tiles_per_ch) * a.fa;
    const int nf = min(a.fa, a.n_frames - f0);
    const int tid = threadIdx.x;

    const int jb = f0 + 1 - a.B;
    const float* qch = a.Q + (size_t)ch * a.nb * a.qs + a.col0;
    // one warp per row, 16-byte loads (col0 and the row stride are multiples of 4 floats)
    const int nc2 = nc >> 1;
    for (int r = tid >> 5; r < nrows; r += 8) {
        const int qr = jb + r - a.j0;
        const bool ok = qr >= 0 && qr < a.nb;
        const float4* src = reinterpret_cast<const float4*>(qch + (size_t)(ok ? qr : 0) * a.qs);
        float2* dst = Qs + r * rs;
        for (int c = tid & 31; c < nc2; c += 32) {
            const float4 v = ok ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            dst[2 * c] = make_float2(v.x, v.y);
            dst[2 * c + 1] = make_float2(v.z, v.w);
        }
    }
    __syncthreads();
    for (int it = tid; it < a.fa * a.nk; it += 256) {
        const int fl = it % a.fa, ki = it / a.fa;
        if (fl >= nf) continue;
        const float2* q = Qs + fl * rs + ki * a.B;
        float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
        for (int b = 0; b < a.B; b += 2) {
            const float2 u = q[b * (rs + 1)], w = q[(b + 1) * (rs + 1)];
            acc0.x += u.x; acc0.y += u.y; acc1.x += w.x; acc1.y += w.y;
        }
        const float re = acc0.x + acc1.x, im = acc0.y + acc1.y;
        mags[fl * a.nk + ki] = cabs_fast(make_float2(re, im)) * a.kw[ki];
    }
    __syncthreads();
    for (int it = tid; it < nf * a.n_tb; it += 256) {
        const int fl = it / a.n_tb, j = it % a.n_tb;
        const int f = f0 + fl;
        const int pos = a.tb_pos[j];
        float val = 0.f;
        if (pos >= 0 && f >= a.first_frame) {
            const float m0 = mags[fl * a.nk + pos], m1 = mags[fl * a.nk + pos + 1];
            const float vi = fmaf(m1 - m0, a.tb_frac[j], m0);
            val = __fdividef(vi * a.wnum, a.wden);
        }
        a.comb_out[((size_t)ch * a.n_frames + f) * a.Tbins + a.tb_idx[j]] = val;
    }
}

inline size_t blockdft_sum_smem_bytes(int nk, int B, int fa) {
    const int nc = nk * B, rs = nc | 1;
    return (size_t)(fa + B - 1) * rs * sizeof(float2) + (size_t)fa * nk * sizeof(float) + 16;
}

inline int blockdft_sum_frames(int nk, int B) {
    int fa = 64;
    while (fa > 16 && blockdft_sum_smem_bytes(nk, B, fa) > 110 * 1024) fa -= 16;   // two CTAs per SM
    return fa;
}

// ---------------------------------------------------------------------------------------------
// Finish of the fused path (blockdft_tc_kernel with X output): |X| * weight and the np.interp segments from
// the complex bins X[ch][bin][frame] the GEMM epilogue assembled.  64 frames per CTA, lanes over frames.
// ---------------------------------------------------------------------------------------------
struct BlockDftFinishArgs {
    const float2* X;           // [n_ch][nkx][n_frames]
    int nkx, xoff, nk;
    const float* kw;           // [nk]
    int n_ch, n_frames, first_frame;
    float* comb_out;
    int Tbins, n_tb;
    const int* tb_idx;
    const int* tb_pos;
    const float* tb_frac;
    float wnum, wden;
};
constexpr int BD_FIN_FRAMES = 64;

__global__ void __launch_bounds__(256)
blockdft_finish_kernel(const __grid_constant__ BlockDftFinishArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* mags = reinterpret_cast<float*>(smem_raw);     // [nk][FRAMES + 1]
    constexpr int MS = BD_FIN_FRAMES + 1;
    const int tiles_per_ch = (a.n_frames + BD_FIN_FRAMES - 1) / BD_FIN_FRAMES;
    const int ch = blockIdx.x / tiles_per_ch;
    const int f0 = (blockIdx.x % tiles_per_ch) * BD_FIN_FRAMES;
    const int nf = min(BD_FIN_FRAMES, a.n_frames - f0);
    const int tid = threadIdx.x;
    const float2* xch = a.X + ((size_t)ch * a.nkx + a.xoff) * a.n_frames + f0;
    for (int it = tid; it < BD_FIN_FRAMES * a.nk; it += 256) {
        const int fl = it % BD_FIN_FRAMES, ki = it / BD_FIN_FRAMES;
        if (fl < nf) mags[ki * MS + fl] = cabs_fast(__ldg(xch + (size_t)ki * a.n_frames + fl)) * a.kw[ki];
    }
    __syncthreads();
    for (int it = tid; it < nf * a.n_tb; it += 256) {
        const int fl = it / a.n_tb, j = it % a.n_tb;
        const int f = f0 + fl;
        const int pos = a.tb_pos[j];
        float val = 0.f;
        if (pos >= 0 && f >= a.first_frame) {
            const float m0 = mags[pos * MS + fl], m1 = mags[(pos + 1) * MS + fl];
            const float vi = fmaf(m1 - m0, a.tb_frac[j], m0);
            val = __fdividef(vi * a.wnum, a.wden);
        }
        a.comb_out[((size_t)ch * a.n_frames + f) * a.Tbins + a.tb_idx[j]] = val;
    }
}

inline size_t blockdft_finish_smem_bytes(int nk) { return (size_t)nk * (BD_FIN_FRAMES + 1) * sizeof(float) + 16; }

}  // namespace o4
