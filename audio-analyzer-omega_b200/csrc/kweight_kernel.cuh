// kweight_kernel.cuh -- per-frame zero-phase K-weighting + mean square -> instantaneous LUFS.
//
// Replaces, for a batch of meter frames:
//   ProfessionalMetering.apply_k_weighting (omega4/panels/professional_meters.py:129-153)
//   the mean-square / LUFS conversion of calculate_lufs (:237-246)
// The reference filters every frame independently with scipy.signal.filtfilt (odd reflection
// padding of 9 samples, steady-state initial conditions zi*x0, forward then backward), twice:
// a Butterworth-2 high-pass at 38 Hz, then a Butterworth-2 high-pass at 1500 Hz blended as
// f + 0.3 (s - f).  No state is carried between frames.
//
// Mapping: ONE WARP PER FRAME, block-parallel linear-recurrence scan.  The padded frame has
// W + 18 = 2066 samples; lane l owns the 65 consecutive samples [65 l, 65 l + 65) in REGISTERS
// (fp64) for all four IIR passes.  Each pass:
//   1. every lane runs the biquad (transposed direct form II, as scipy's lfilter) over its chunk
//      from zero state  -> zero-state outputs + end state e_l
//   2. the true chunk-entry states follow v_l = Phi v_{l-1} + e_l (Phi = A^65) -> 5-step
//      Kogge-Stone scan with shuffles, multipliers Phi^(2^j) precomputed on the host
//   3. every lane adds the homogeneous response g[i] . s_in to its outputs
// Backward passes run the same code with lane order and sample order reversed.  The 14 slack
// positions of lane 31 are zeroed before a backward pass and the steady-state start
// zi * y[last] is moved 14 steps back in time (s_virt = A^-14 zi y_last) so that all lanes stay
// uniform.  fp64 state is required: the 38 Hz poles sit at radius 0.9965 and fp32 coefficients
// alone would move the cut-off by ~0.2 % ("IIR precision").
#pragma once
#include <cuda_runtime.h>

namespace o4 {

constexpr int KW_W = 2048;          // meter frame length (FFT_SIZE_BASE)
constexpr int KW_PAD = 9;           // filtfilt padlen = 3 * max(len(a), len(b))
constexpr int KW_L = 65;            // samples per lane
constexpr int KW_EXT = KW_W + 2 * KW_PAD;       // 2066
constexpr int KW_SLACK = 32 * KW_L - KW_EXT;    // 14
constexpr int KW_LAST = KW_L - 1 - KW_SLACK;    // 50: local index of ext[2065] on lane 31

struct KwBiquad {
    double b0, b1, b2, a1, a2;
    double zi0, zi1;                // lfilter_zi
    double phi[5][4];               // (A^65)^(2^j), row major [[p00,p01],[p10,p11]]
    double ainv[4];                 // A^-14
    double g[KW_L][2];              // g[i] = first row of A^i  (homogeneous output response)
};

struct KweightArgs {
    const void* x;                  // float samples (hop mode) or double frames (frames mode)
    int x_is_f64;
    long long ch_stride;
    long long frame_stride;
    long long frame_off0;
    int n_ch;
    int n_frames;
    int first_frame;
    int frames_per_warp;
    const double* hann;             // [W] float64 Hann (hop mode) or nullptr (frames already windowed)
    double* lufs_out;               // [n_ch][n_frames] instantaneous LUFS
    double* weighted_out;           // [n_ch][n_frames][W] K-weighted frame or nullptr
    KwBiquad f[2];                  // 0: 38 Hz high-pass, 1: 1500 Hz "shelf" high-pass
};

__device__ __forceinline__ double2 mat2_apply(const double* m, double2 s) {
    return make_double2(fma(m[0], s.x, m[1] * s.y), fma(m[2], s.x, m[3] * s.y));
}

// one lfilter pass over the warp's 2066-sample sequence held as r[65] per lane.
template <bool BACKWARD>
__device__ __forceinline__ void kw_pass(double (&r)[KW_L], const KwBiquad& c, int lane) {
    // scan position of this lane: 0 is processed first
    const int pos = BACKWARD ? 31 - lane : lane;
    // the sample whose value scales the steady-state initial condition
    double x0 = BACKWARD ? __shfl_sync(0xffffffffu, r[KW_LAST], 31) : __shfl_sync(0xffffffffu, r[0], 0);
    if (BACKWARD && lane == 31) {
#pragma unroll
        for (int i = KW_LAST + 1; i < KW_L; ++i) r[i] = 0.0;
    }
    double2 s_init = make_double2(c.zi0 * x0, c.zi1 * x0);
    if (BACKWARD) s_init = mat2_apply(c.ainv, s_init);

    // 1. zero-state sweep
    double z1 = 0.0, z2 = 0.0;
#pragma unroll
    for (int n = 0; n < KW_L; ++n) {
        const int i = BACKWARD ? KW_L - 1 - n : n;
        double x = r[i];
        double y = fma(c.b0, x, z1);
        z1 = fma(-c.a1, y, fma(c.b1, x, z2));
        z2 = fma(-c.a2, y, c.b2 * x);
        r[i] = y;
    }
    // 2. scan of chunk end states
    double2 v = make_double2(z1, z2);
    if (pos == 0) {
        double2 q = mat2_apply(c.phi[0], s_init);
        v.x += q.x; v.y += q.y;
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int d = 1 << j;
        double rx = BACKWARD ? __shfl_down_sync(0xffffffffu, v.x, d) : __shfl_up_sync(0xffffffffu, v.x, d);
        double ry = BACKWARD ? __shfl_down_sync(0xffffffffu, v.y, d) : __shfl_up_sync(0xffffffffu, v.y, d);
        if (pos >= d) {
            double2 q = mat2_apply(c.phi[j], make_double2(rx, ry));
            v.x += q.x; v.y += q.y;
        }
    }
    double sx = BACKWARD ? __shfl_down_sync(0xffffffffu, v.x, 1) : __shfl_up_sync(0xffffffffu, v.x, 1);
    double sy = BACKWARD ? __shfl_down_sync(0xffffffffu, v.y, 1) : __shfl_up_sync(0xffffffffu, v.y, 1);
    if (pos == 0) { sx = s_init.x; sy = s_init.y; }
    // 3. homogeneous correction
#pragma unroll
    for (int n = 0; n < KW_L; ++n) {
        const int i = BACKWARD ? KW_L - 1 - n : n;
        r[i] = fma(c.g[n][0], sx, fma(c.g[n][1], sy, r[i]));
    }
}

// odd reflection padding of the frame held at ext positions [9, 2057): lane 0 owns ext[0..9),
// lane 31 owns ext[2057..2066) at local 42..50.
__device__ __forceinline__ void kw_odd_pad(double (&r)[KW_L], int lane) {
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < KW_PAD; ++j) r[j] = 2.0 * r[KW_PAD] - r[2 * KW_PAD - j];
    }
    if (lane == 31) {
        constexpr int E = KW_LAST - KW_PAD;     // 41: local index of the frame's last sample
#pragma unroll
        for (int j = 0; j < KW_PAD; ++j) r[E + 1 + j] = 2.0 * r[E] - r[E - 1 - j];
    }
}

constexpr int KW_WARPS = 4;

__global__ void __launch_bounds__(KW_WARPS * 32, 2)
kweight_kernel(const __grid_constant__ KweightArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* hann_s = reinterpret_cast<double*>(smem_raw);                          // [W]
    float* stage_all = reinterpret_cast<float*>(hann_s + KW_W);                     // [KW_WARPS][32*65]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* stg = stage_all + warp * (32 * KW_L);

    if (a.hann) {
        for (int i = threadIdx.x; i < KW_W; i += blockDim.x) hann_s[i] = a.hann[i];
    }
    __syncthreads();

    const int frames_per_cta = a.frames_per_warp * KW_WARPS;
    const int tiles_per_ch = (a.n_frames + frames_per_cta - 1) / frames_per_cta;
    const int ch = blockIdx.x / tiles_per_ch;
    const int tile = blockIdx.x % tiles_per_ch;

    for (int it = 0; it < a.frames_per_warp; ++it) {
        const int f = tile * frames_per_cta + it * KW_WARPS + warp;
        if (f >= a.n_frames || f < a.first_frame) continue;           // warp-uniform
        const long long off = (long long)ch * a.ch_stride + a.frame_off0 + (long long)f * a.frame_stride;

        double r[KW_L];
        double sumsq = 0.0;
        if (!a.x_is_f64) {
            // coalesced copy of the frame into shared memory, then strided (conflict free) reads
            const float4* px = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.x) + off);
            float4* s4 = reinterpret_cast<float4*>(stg);
#pragma unroll
            for (int j = 0; j < KW_W / 4 / 32; ++j) s4[lane + 32 * j] = __ldg(px + lane + 32 * j);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < KW_L; ++i) {
                int n = KW_L * lane + i - KW_PAD;                      // frame sample index
                double xv = 0.0;
                if (n >= 0 && n < KW_W) {
                    xv = (double)stg[n];
                    if (a.hann) xv *= hann_s[n];
                    sumsq = fma(xv, xv, sumsq);
                }
                r[i] = xv;
            }
            __syncwarp();
        } else {
            const double* px = reinterpret_cast<const double*>(a.x) + off;
#pragma unroll
            for (int i = 0; i < KW_L; ++i) {
                int n = KW_L * lane + i - KW_PAD;
                double xv = 0.0;
                if (n >= 0 && n < KW_W) {
                    xv = px[n];
                    if (a.hann) xv *= hann_s[n];
                    sumsq = fma(xv, xv, sumsq);
                }
                r[i] = xv;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
        const bool gated = sqrt(sumsq / (double)KW_W) < 1e-6;          // professional_meters.py:132-134

        double ms = 0.0;
        if (!gated) {                                                  // warp-uniform
            kw_odd_pad(r, lane);
            kw_pass<false>(r, a.f[0], lane);
            kw_pass<true>(r, a.f[0], lane);
            // stash f (first filtfilt output) as fp32; 65-word lane stride is bank-conflict free
#pragma unroll
            for (int i = 0; i < KW_L; ++i) stg[KW_L * lane + i] = (float)r[i];
            kw_odd_pad(r, lane);
            kw_pass<false>(r, a.f[1], lane);
            kw_pass<true>(r, a.f[1], lane);
            double acc = 0.0;
            double* wrow = a.weighted_out ? a.weighted_out + ((size_t)ch * a.n_frames + f) * KW_W : nullptr;
#pragma unroll
            for (int i = 0; i < KW_L; ++i) {
                int n = KW_L * lane + i - KW_PAD;
                if (n >= 0 && n < KW_W) {
                    double fv = (double)stg[KW_L * lane + i];
                    double w = fma(r[i] - fv, 0.3, fv);                // f + (s - f) * 0.3
                    acc = fma(w, w, acc);
                    if (wrow) wrow[n] = w;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            ms = acc / (double)KW_W;
            __syncwarp();
        } else if (a.weighted_out) {
            double* wrow = a.weighted_out + ((size_t)ch * a.n_frames + f) * KW_W;
            for (int n = lane; n < KW_W; n += 32) wrow[n] = 0.0;
        }
        if (lane == 0)
            a.lufs_out[(size_t)ch * a.n_frames + f] = (ms > 1e-10) ? (-0.691 + 10.0 * log10(ms)) : -100.0;
    }
}

inline size_t kweight_smem_bytes() {
    return (size_t)KW_W * sizeof(double) + (size_t)KW_WARPS * 32 * KW_L * sizeof(float);
}

}  // namespace o4
