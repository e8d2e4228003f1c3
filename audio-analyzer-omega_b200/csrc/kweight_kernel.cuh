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
//   1. every lane runs the biquad over its chunk from zero OUTPUT state (the two previous inputs
//      come from the neighbouring lane) -> zero-state outputs + end state e_l = (y[64], y[63]).
//      The recursion is written in direct form I, y[n] = (u[n] - a2 y[n-2]) - a1 y[n-1], so that
//      only ONE fp64 FMA per sample sits on the dependency chain (scipy's transposed form II has
//      two; ncu showed the kernel bound by exactly that fixed-latency chain).  scipy's initial
//      state zi*x0 is converted once into the equivalent (y[-1], y[-2]).
//   2. the true chunk-entry states follow v_l = Phi v_{l-1} + e_l (Phi = C^65, C the companion
//      matrix) -> 5-step Kogge-Stone scan with shuffles, multipliers Phi^(2^j) from the host
//   3. every lane adds the homogeneous response g[i] . s_in to its outputs
// Backward passes run the same code with lane order and sample order reversed.  The 14 slack
// positions of lane 31 are zeroed before a backward pass and the steady-state start is moved 14
// steps back in time (s_virt = C^-14 s_init) so that all lanes stay uniform.  fp64 state is required: the 38 Hz poles sit at radius 0.9965 and fp32 coefficients
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
// every lane's 65 samples are swept as 4 independent sub-chunks of 17+16+16+16 samples: four
// interleaved dependency chains hide the fp64 FMA latency that a single chain exposes
constexpr int KW_NSUB = 4;
constexpr int KW_SUBMAX = 17;
__host__ __device__ constexpr int kw_off(int j) { return j == 0 ? 0 : 17 + 16 * (j - 1); }   // 0,17,33,49,65

template <typename T>
struct KwBiquadT {
    T b0, b1, b2, a1, a2;
    T yi0, yi1;                     // (y[-1], y[-2]) per unit x0, equivalent to scipy's lfilter_zi state
    T phi[5][4];                    // (C^65)^(2^j), row major [[p00,p01],[p10,p11]], C = [[-a1,-a2],[1,0]]
    T cinv_f[4];                    // C^-(9 - pad): forward start moved back over the unused leading slots
    T cinv_b[4];                    // C^-(14 + 9 - pad): backward start moved back over the trailing slack
    int pad;                        // filtfilt padlen = 3 * max(len(a), len(b)): 9 (biquad) or 6 (first order)
    int _align;
    T c16[4], c17[4];               // C^16, C^17: transitions over one sub-chunk
    T g[KW_SUBMAX][2];              // g[i] = first row of C^(i+1)  (homogeneous output response)
};
using KwBiquad = KwBiquadT<double>;
using KwBiquadF = KwBiquadT<float>;   // float32 mirror, used for well-conditioned sections only

template <typename T> struct KwVec2;
template <> struct KwVec2<double> { using type = double2; };
template <> struct KwVec2<float> { using type = float2; };
__device__ __forceinline__ double2 kw_make2(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ float2 kw_make2(float x, float y) { return make_float2(x, y); }

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
    // weighting program (professional_meters.py:129-229): n_sec zero-phase sections applied in cascade.
    //   K: 2 sections, blend = 1: out = f0 + 0.3 (f1 - f0), f1 = section 1 applied to f0   (:137-151)
    //   A: 4 sections, gain 2.5 (:166-190);  C: 2 sections (:205-216);  Z: 0 sections, no gate (:228-229)
    int n_sec;
    int blend;
    int rms_gate;
    int _align;
    double gain;
    // sections [f32_from, n_sec) run in float32: their poles are far from the unit circle (|z| < 0.95),
    // where float32 state costs ~1e-6 relative error -- 1e-5 LU against the 0.01 LU parity bar -- and
    // half the pipe time of float64.  Sections before f32_from (the 38 Hz / 20.6 Hz high-passes with
    // poles at radius 0.997) keep float64 state.  Kernels that return the weighted frame use float64
    // throughout (f32_from = n_sec).
    int f32_from;
    int _align2;
    KwBiquad f[4];
    KwBiquadF ff[4];
};

template <typename T>
__device__ __forceinline__ typename KwVec2<T>::type mat2_apply(const T* m, typename KwVec2<T>::type s) {
    return kw_make2(fma(m[0], s.x, m[1] * s.y), fma(m[2], s.x, m[3] * s.y));
}

// one lfilter pass over the warp's 2066-sample sequence held as r[65] per lane.
template <bool BACKWARD, typename T>
__device__ __forceinline__ void kw_pass(T (&r)[KW_L], const KwBiquadT<T>& c, int lane) {
    using V2 = typename KwVec2<T>::type;
    const T zero = (T)0;
    // scan position of this lane: 0 is processed first
    const int pos = BACKWARD ? 31 - lane : lane;
    // the sample whose value scales the steady-state initial condition: first / last sample of the
    // padded sequence, which starts 9 - pad slots into lane 0 and ends 9 - pad slots before KW_LAST
    const bool short_pad = c.pad < KW_PAD;                           // pad 6: three unused slots at either end
    constexpr int SH = KW_PAD - 6;
    T x0 = BACKWARD ? __shfl_sync(0xffffffffu, short_pad ? r[KW_LAST - SH] : r[KW_LAST], 31)
                         : __shfl_sync(0xffffffffu, short_pad ? r[SH] : r[0], 0);
    if (BACKWARD && lane == 31) {
#pragma unroll
        for (int i = KW_LAST + 1; i < KW_L; ++i) r[i] = zero;
        if (short_pad) {
#pragma unroll
            for (int i = KW_LAST - SH + 1; i <= KW_LAST; ++i) r[i] = zero;
        }
    }
    V2 s_init = kw_make2(c.yi0 * x0, c.yi1 * x0);
    s_init = mat2_apply<T>(BACKWARD ? c.cinv_b : c.cinv_f, s_init);

    // sub-chunk m (processing order) = index range j: forward j = m, backward j = 3 - m
    // length of sub-chunk m in processing order: forward 17,16,16,16 ; backward 16,16,16,17
    T xa[KW_NSUB], xb[KW_NSUB];      // the two inputs preceding each sub-chunk (processing order)
    xa[0] = BACKWARD ? __shfl_down_sync(0xffffffffu, r[0], 1) : __shfl_up_sync(0xffffffffu, r[KW_L - 1], 1);
    xb[0] = BACKWARD ? __shfl_down_sync(0xffffffffu, r[1], 1) : __shfl_up_sync(0xffffffffu, r[KW_L - 2], 1);
    if (pos == 0) { xa[0] = zero; xb[0] = zero; }
#pragma unroll
    for (int m = 1; m < KW_NSUB; ++m) {
        const int j = BACKWARD ? KW_NSUB - 1 - m : m;
        xa[m] = BACKWARD ? r[kw_off(j + 1)] : r[kw_off(j) - 1];
        xb[m] = BACKWARD ? r[kw_off(j + 1) + 1] : r[kw_off(j) - 2];
    }
    // 1. zero-state sweeps, direct form I, four independent chains
    T y1[KW_NSUB], y2[KW_NSUB];
#pragma unroll
    for (int m = 0; m < KW_NSUB; ++m) { y1[m] = zero; y2[m] = zero; }
#pragma unroll
    for (int n = 0; n < KW_SUBMAX; ++n) {
#pragma unroll
        for (int m = 0; m < KW_NSUB; ++m) {
            const int j = BACKWARD ? KW_NSUB - 1 - m : m;
            const int len = kw_off(j + 1) - kw_off(j);
            if (n < len) {
                const int i = BACKWARD ? kw_off(j + 1) - 1 - n : kw_off(j) + n;
                const T x = r[i];
                const T u = fma(c.b0, x, fma(c.b1, xa[m], c.b2 * xb[m]));
                const T t = fma(-c.a2, y2[m], u);
                const T y = fma(-c.a1, y1[m], t);
                xb[m] = xa[m]; xa[m] = x;
                y2[m] = y1[m]; y1[m] = y;
                r[i] = y;
            }
        }
    }
    // lane aggregate: zero-state end state of the whole 65-sample chunk
    V2 v = kw_make2(y1[0], y2[0]);
#pragma unroll
    for (int m = 1; m < KW_NSUB; ++m) {
        const int j = BACKWARD ? KW_NSUB - 1 - m : m;
        const T* cm = (kw_off(j + 1) - kw_off(j) == 17) ? c.c17 : c.c16;
        V2 q = mat2_apply<T>(cm, v);
        v = kw_make2(q.x + y1[m], q.y + y2[m]);
    }
    // 2. scan of chunk end states across lanes
    if (pos == 0) {
        V2 q = mat2_apply<T>(c.phi[0], s_init);
        v.x += q.x; v.y += q.y;
    }
#pragma unroll
    for (int jj = 0; jj < 5; ++jj) {
        const int d = 1 << jj;
        T rx = BACKWARD ? __shfl_down_sync(0xffffffffu, v.x, d) : __shfl_up_sync(0xffffffffu, v.x, d);
        T ry = BACKWARD ? __shfl_down_sync(0xffffffffu, v.y, d) : __shfl_up_sync(0xffffffffu, v.y, d);
        if (pos >= d) {
            V2 q = mat2_apply<T>(c.phi[jj], kw_make2(rx, ry));
            v.x += q.x; v.y += q.y;
        }
    }
    V2 sin[KW_NSUB];
    sin[0].x = BACKWARD ? __shfl_down_sync(0xffffffffu, v.x, 1) : __shfl_up_sync(0xffffffffu, v.x, 1);
    sin[0].y = BACKWARD ? __shfl_down_sync(0xffffffffu, v.y, 1) : __shfl_up_sync(0xffffffffu, v.y, 1);
    if (pos == 0) sin[0] = s_init;
    // entry state of every sub-chunk: s_m = C^len(m-1) s_(m-1) + e_(m-1)
#pragma unroll
    for (int m = 1; m < KW_NSUB; ++m) {
        const int jp = BACKWARD ? KW_NSUB - m : m - 1;              // index range of sub-chunk m-1
        const T* cm = (kw_off(jp + 1) - kw_off(jp) == 17) ? c.c17 : c.c16;
        V2 q = mat2_apply<T>(cm, sin[m - 1]);
        sin[m] = kw_make2(q.x + y1[m - 1], q.y + y2[m - 1]);
    }
    // 3. homogeneous correction
#pragma unroll
    for (int n = 0; n < KW_SUBMAX; ++n) {
#pragma unroll
        for (int m = 0; m < KW_NSUB; ++m) {
            const int j = BACKWARD ? KW_NSUB - 1 - m : m;
            const int len = kw_off(j + 1) - kw_off(j);
            if (n < len) {
                const int i = BACKWARD ? kw_off(j + 1) - 1 - n : kw_off(j) + n;
                r[i] = fma(c.g[n][0], sin[m].x, fma(c.g[n][1], sin[m].y, r[i]));
            }
        }
    }
}

// odd reflection padding of the frame held at ext positions [9, 2057): lane 0 owns ext[0..9),
// lane 31 owns ext[2057..2066) at local 42..50.
template <typename T>
__device__ __forceinline__ void kw_odd_pad(T (&r)[KW_L], int lane, int pad) {
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < KW_PAD; ++j) r[j] = (j >= KW_PAD - pad) ? (T)2 * r[KW_PAD] - r[2 * KW_PAD - j] : (T)0;
    }
    if (lane == 31) {
        constexpr int E = KW_LAST - KW_PAD;     // 41: local index of the frame's last sample
#pragma unroll
        for (int j = 0; j < KW_PAD; ++j) r[E + 1 + j] = (j < pad) ? (T)2 * r[E] - r[E - 1 - j] : (T)0;
    }
}

constexpr int KW_WARPS = 4;
#ifndef KW_MIN_BLOCKS
#define KW_MIN_BLOCKS 2
#endif
// per-warp staging strip in "ext" coordinates shifted by 3 floats so that the frame's first sample
// (ext position 9) lands on a 16-byte boundary: index = ext position + 3
constexpr int KW_STG_SHIFT = 3;
constexpr int KW_STG = 32 * KW_L + 16;          // 2096 floats per warp

template <typename T>
__device__ __forceinline__ void kw_zero_pads(T (&r)[KW_L], int lane) {
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < KW_PAD; ++i) r[i] = (T)0;
    }
    if (lane == 31) {
#pragma unroll
        for (int i = KW_LAST - KW_PAD + 1; i < KW_L; ++i) r[i] = (T)0;
    }
}

// F64_FRAMES: input is explicit float64 frames (streaming shim), else float32 samples + Hann table.
// WEIGHTED: also store the K-weighted frame (ProfessionalMetering.apply_k_weighting).
template <bool F64_FRAMES, bool WEIGHTED>
__global__ void __launch_bounds__(KW_WARPS * 32, KW_MIN_BLOCKS)
kweight_kernel(const __grid_constant__ KweightArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* hann_x = reinterpret_cast<double*>(smem_raw);                          // [32*65] ext layout, 0 at pads
    float* stage_all = reinterpret_cast<float*>(hann_x + 32 * KW_L);               // [KW_WARPS][KW_STG]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* stg = stage_all + warp * KW_STG;

    for (int p = threadIdx.x; p < 32 * KW_L; p += blockDim.x) {
        const int n = p - KW_PAD;
        hann_x[p] = (n >= 0 && n < KW_W) ? (a.hann ? a.hann[n] : 1.0) : 0.0;
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
        if (!F64_FRAMES) {
            // coalesced copy of the frame into the staging strip (pads zeroed), then conflict-free
            // strided reads: no per-element predicates anywhere
            const float4* px = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.x) + off);
            float4* s4 = reinterpret_cast<float4*>(stg + KW_PAD + KW_STG_SHIFT);
            float4 tmp[KW_W / 4 / 32];
#pragma unroll
            for (int j = 0; j < KW_W / 4 / 32; ++j) tmp[j] = __ldg(px + lane + 32 * j);
            if (lane < KW_PAD) stg[KW_STG_SHIFT + lane] = 0.f;
            if (lane < KW_SLACK + KW_PAD + KW_STG_SHIFT) stg[KW_STG_SHIFT + KW_PAD + KW_W + lane] = 0.f;
#pragma unroll
            for (int j = 0; j < KW_W / 4 / 32; ++j) s4[lane + 32 * j] = tmp[j];
            __syncwarp();
            const float* sl = stg + KW_STG_SHIFT + KW_L * lane;
            const double* hl = hann_x + KW_L * lane;
#pragma unroll
            for (int i = 0; i < KW_L; ++i) {
                const double xv = (double)sl[i] * hl[i];
                sumsq = fma(xv, xv, sumsq);
                r[i] = xv;
            }
            __syncwarp();
        } else {
            const double* px = reinterpret_cast<const double*>(a.x) + off;
#pragma unroll
            for (int i = 0; i < KW_L; ++i) {
                const int n = KW_L * lane + i - KW_PAD;
                double xv = 0.0;
                if (n >= 0 && n < KW_W) xv = px[n] * hann_x[KW_L * lane + i];
                sumsq = fma(xv, xv, sumsq);
                r[i] = xv;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
        const bool gated = a.rms_gate && sqrt(sumsq / (double)KW_W) < 1e-6;   // professional_meters.py:132-134

        double ms = 0.0;
        if (!gated) {                                                  // warp-uniform
            float* fl = stg + KW_STG_SHIFT + KW_L * lane;
            // the filtfilt calls of one precision share one code body (loop not unrolled: halves the
            // instruction footprint -- ncu showed 14 % instruction-fetch stalls with both inlined)
            const int n64 = WEIGHTED ? a.n_sec : min(a.f32_from, a.n_sec);
#pragma unroll 1
            for (int fi = 0; fi < n64; ++fi) {
                kw_odd_pad(r, lane, a.f[fi].pad);
                kw_pass<false, double>(r, a.f[fi], lane);
                kw_pass<true, double>(r, a.f[fi], lane);
                kw_zero_pads(r, lane);
                if (fi == 0 && a.blend) {
                    // stash f (first filtfilt output) as fp32 with zeros at the pad positions;
                    // the 65-word lane stride is bank-conflict free
#pragma unroll
                    for (int i = 0; i < KW_L; ++i) fl[i] = (float)r[i];
                }
            }
            double acc = 0.0;
            if (!WEIGHTED && n64 < a.n_sec) {
                // well-conditioned tail of the cascade in float32
                float rf[KW_L];
#pragma unroll
                for (int i = 0; i < KW_L; ++i) rf[i] = (float)r[i];
#pragma unroll 1
                for (int fi = n64; fi < a.n_sec; ++fi) {
                    kw_odd_pad(rf, lane, a.ff[fi].pad);
                    kw_pass<false, float>(rf, a.ff[fi], lane);
                    kw_pass<true, float>(rf, a.ff[fi], lane);
                    kw_zero_pads(rf, lane);
                }
                const float gain = (float)a.gain;
                float accf = 0.f;
#pragma unroll
                for (int i = 0; i < KW_L; ++i) {
                    float w;
                    if (a.blend) {
                        const float fv = fl[i];
                        w = fmaf(rf[i] - fv, 0.3f, fv);                // f + (s - f) * 0.3 ; 0 at the pads
                    } else {
                        w = rf[i] * gain;
                    }
                    accf = fmaf(w, w, accf);
                }
                acc = (double)accf;
            } else {
#pragma unroll
                for (int i = 0; i < KW_L; ++i) {
                    double w;
                    if (a.blend) {
                        const double fv = (double)fl[i];
                        w = fma(r[i] - fv, 0.3, fv);                   // f + (s - f) * 0.3 ; 0 at the pads
                    } else {
                        w = r[i] * a.gain;                             // cascade (A: x 2.5, C / Z: x 1)
                    }
                    acc = fma(w, w, acc);
                    if (WEIGHTED) r[i] = w;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            ms = acc / (double)KW_W;
            if (WEIGHTED && a.weighted_out) {
                double* wrow = a.weighted_out + ((size_t)ch * a.n_frames + f) * KW_W;
#pragma unroll
                for (int i = 0; i < KW_L; ++i) {
                    const int n = KW_L * lane + i - KW_PAD;
                    if (n >= 0 && n < KW_W) wrow[n] = r[i];
                }
            }
            __syncwarp();
        } else if (WEIGHTED && a.weighted_out) {
            double* wrow = a.weighted_out + ((size_t)ch * a.n_frames + f) * KW_W;
            for (int n = lane; n < KW_W; n += 32) wrow[n] = 0.0;
        }
        if (lane == 0)
            a.lufs_out[(size_t)ch * a.n_frames + f] = (ms > 1e-10) ? (-0.691 + 10.0 * log10(ms)) : -100.0;
    }
}

inline size_t kweight_smem_bytes() {
    return (size_t)32 * KW_L * sizeof(double) + (size_t)KW_WARPS * KW_STG * sizeof(float);
}

}  // namespace o4
