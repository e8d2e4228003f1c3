// fft_core.cuh -- shared-memory Stockham FFT building blocks for sm_100a.
//
// Layout of one complex transform of M = 2^LOG2M points ("sub-FFT"):
//   * M/16 threads cooperate, every thread always owns 16 points ("16 points per thread"
//     invariant), so window values and stage twiddles are per-thread constants that live in
//     registers across all frames a CTA processes.
//   * Decimation-in-frequency Stockham autosort, radix 16 while the sub-length n >= 16, then
//     one radix-2/4/8 tail stage.  For a stage (n, s):
//         y[q + s*(16p + k)] = W_n^{pk} * sum_j x[q + s*(p + j*n/16)] * W_16^{jk}
//     with butterfly index t = q + s*p == the thread index, so EVERY stage reads x[t + j*M/16]
//     (conflict free) and the result ends up in natural order.
//   * Shared-memory index padding pad(i) = i + (i >> 4) (float2 units) makes the strided
//     writes of the first two stages conflict free (17t + k, q + 272p + 17k).
//   * Stage twiddles W_M^{e}: 4 table loads per stage (powers 1,2,4,8), the other 11 powers are
//     products of depth <= 3 -- keeps fp32 twiddle error at a few ulp without table traffic.
//
// A real N-point transform is done as an N/2-point complex transform of (even, odd) pairs plus
// an "untangle" epilogue (see rfft_pair()).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace o4 {

// Complex arithmetic on (re, im) register pairs.  Additions, the sqrt(1/2) rotations and element-wise products are
// issued as packed float32 instructions (add.f32x2 / mul.f32x2, sm_100: FADD2 / FMUL2): the same IEEE operations bit
// for bit, one issue slot instead of two (the FMA pipe's lane rate is unchanged, tools/micro/f32x2_pipes.cu), which is
// what these kernels are short of next to their shared-memory traffic: N = 2048 -8 %, N = 1024 -5 % (r02m).  ptxas
// folds the (a.y, -a.x) swaps and half negations into operand selectors (.LO_HI, .NP), no moves.  The general complex
// product stays scalar: its packed form (FMUL2 + FFMA2 with broadcast selectors) is also bit-identical, but pairs up
// registers until ptxas pays for it in moves -- measured slower (N = 1024: 3.83 -> 4.10 ms at 256 streams x 30 s).
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a * (r, -r) and a * (-r, -r), r = sqrt(1/2): one FADD2 + one FMUL2 each
__device__ __forceinline__ float2 rot_m45(float2 a, float r) {
    return __fmul2_rn(__fadd2_rn(a, make_float2(a.y, -a.x)), make_float2(r, r));
}
__device__ __forceinline__ float2 rot_m135(float2 a, float r) {
    return __fmul2_rn(__fadd2_rn(make_float2(a.y, -a.x), make_float2(-a.x, -a.y)), make_float2(r, r));
}
__device__ __forceinline__ float2 cmul_elem(float2 a, float2 b) { return __fmul2_rn(a, b); }      // (a.x b.x, a.y b.y)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// multiply by -i : (x, y) -> (y, -x)
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

__host__ __device__ __forceinline__ constexpr int padi(int i) { return i + (i >> 4); }

// forward radix-4 butterfly, outputs in natural order
__device__ __forceinline__ void bf4(float2& x0, float2& x1, float2& x2, float2& x3) {
    float2 a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = mul_mi(csub(x1, x3));
    x0 = cadd(a, c); x1 = cadd(b, d); x2 = csub(a, c); x3 = csub(b, d);
}
__device__ __forceinline__ void bf2(float2& x0, float2& x1) {
    float2 a = cadd(x0, x1), b = csub(x0, x1);
    x0 = a; x1 = b;
}

// forward radix-8 butterfly (in place, natural order out)
__device__ __forceinline__ void bf8(float2* v) {
    const float r = 0.70710678118654752440f;
    // 8 = 2 x 4 : a in {0,1}, j = a + 2b
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    bf4(e0, e1, e2, e3);
    bf4(o0, o1, o2, o3);
    // twiddle o_c by W8^c
    o1 = rot_m45(o1, r);                                         // * (r, -r)
    o2 = mul_mi(o2);                                             // * -i
    o3 = rot_m135(o3, r);                                        // * (-r, -r)
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// forward radix-16 butterfly: out[k] = sum_j in[j] W16^{jk}, in place, natural order out.
__device__ __forceinline__ void bf16pt(float2* v) {
    const float c1 = 0.92387953251128675613f;   // cos(pi/8)
    const float s1 = 0.38268343236508977173f;   // sin(pi/8)
    const float r = 0.70710678118654752440f;
    // step 1: for a in 0..3 radix-4 over j = a + 4b  -> t[a][c] stored at v[a + 4c]
#pragma unroll
    for (int a = 0; a < 4; ++a) bf4(v[a], v[a + 4], v[a + 8], v[a + 12]);
    // twiddle t[a][c] *= W16^{a*c}   (v index a + 4c)
    // a=1: c=1 -> W^1, c=2 -> W^2, c=3 -> W^3
    v[1 + 4]  = cmul(v[1 + 4],  make_float2(c1, -s1));
    v[1 + 8]  = rot_m45(v[1 + 8], r);
    v[1 + 12] = cmul(v[1 + 12], make_float2(s1, -c1));
    // a=2: c=1 -> W^2, c=2 -> W^4 = -i, c=3 -> W^6 = (-r,-r)
    v[2 + 4]  = rot_m45(v[2 + 4], r);
    v[2 + 8]  = mul_mi(v[2 + 8]);
    v[2 + 12] = rot_m135(v[2 + 12], r);
    // a=3: c=1 -> W^3, c=2 -> W^6, c=3 -> W^9 = -W^1
    v[3 + 4]  = cmul(v[3 + 4],  make_float2(s1, -c1));
    v[3 + 8]  = rot_m135(v[3 + 8], r);
    v[3 + 12] = cmul(v[3 + 12], make_float2(-c1, s1));
    // step 2: for c: radix-4 over a -> out[c + 4d] ; t[a][c] sits at v[a + 4c]
#pragma unroll
    for (int c = 0; c < 4; ++c) bf4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    // now v[4c + d] holds out[c + 4d]  -> transpose the 4x4 index to natural order
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) { float2 tmp = v[4 * c + d]; v[4 * c + d] = v[4 * d + c]; v[4 * d + c] = tmp; }
}

// out[k] *= w^k, k = 1..15, from the four exact table powers w1,w2,w4,w8.
struct Tw4 { float2 w1, w2, w4, w8; };

__device__ __forceinline__ void apply_twiddles(float2* v, const Tw4& t) {
    float2 w3 = cmul(t.w1, t.w2);
    float2 w5 = cmul(t.w1, t.w4);
    float2 w6 = cmul(t.w2, t.w4);
    float2 w7 = cmul(w3, t.w4);
    v[1] = cmul(v[1], t.w1);
    v[2] = cmul(v[2], t.w2);
    v[3] = cmul(v[3], w3);
    v[4] = cmul(v[4], t.w4);
    v[5] = cmul(v[5], w5);
    v[6] = cmul(v[6], w6);
    v[7] = cmul(v[7], w7);
    v[8] = cmul(v[8], t.w8);
    v[9] = cmul(v[9], cmul(t.w1, t.w8));
    v[10] = cmul(v[10], cmul(t.w2, t.w8));
    v[11] = cmul(v[11], cmul(w3, t.w8));
    v[12] = cmul(v[12], cmul(t.w4, t.w8));
    v[13] = cmul(v[13], cmul(w5, t.w8));
    v[14] = cmul(v[14], cmul(w6, t.w8));
    v[15] = cmul(v[15], cmul(w7, t.w8));
}

// all 15 powers precomputed (30 registers per stage instead of 8 + 44 instructions per use)
struct Tw15 { float2 w[15]; };

__device__ __forceinline__ void apply_twiddles(float2* v, const Tw15& t) {
#pragma unroll
    for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], t.w[k - 1]);
}

__device__ __forceinline__ void expand_twiddles(const Tw4& t, Tw15& o) {
    const float2 w3 = cmul(t.w1, t.w2), w5 = cmul(t.w1, t.w4), w6 = cmul(t.w2, t.w4), w7 = cmul(w3, t.w4);
    o.w[0] = t.w1; o.w[1] = t.w2; o.w[2] = w3; o.w[3] = t.w4; o.w[4] = w5; o.w[5] = w6; o.w[6] = w7; o.w[7] = t.w8;
    o.w[8] = cmul(t.w1, t.w8); o.w[9] = cmul(t.w2, t.w8); o.w[10] = cmul(w3, t.w8); o.w[11] = cmul(t.w4, t.w8);
    o.w[12] = cmul(w5, t.w8); o.w[13] = cmul(w6, t.w8); o.w[14] = cmul(w7, t.w8);
}

// ---------------------------------------------------------------------------------------------
// Compile-time description of one transform size.
// ---------------------------------------------------------------------------------------------
template <int LOG2M>
struct FftShape {
    static constexpr int M = 1 << LOG2M;          // complex points
    static constexpr int TPF = M / 16;            // threads per sub-FFT
    static constexpr int NT = (TPF > 256) ? TPF : 256;   // CTA threads
    static constexpr int CONC = NT / TPF;         // sub-FFTs processed concurrently by a CTA
    static constexpr int BUF = M + M / 16;        // padded float2 entries per sub-FFT buffer
    static constexpr int NTW = (LOG2M - 1) / 4;   // number of twiddled radix-16 stages (n >= 32)
    static constexpr int TAIL = M >> (4 * (LOG2M / 4));          // 1, 2, 4 or 8
    static constexpr bool LAST16 = (LOG2M % 4 == 0);             // final untwiddled radix-16 stage
    // two ping-pong buffers per sub-FFT, except for the largest size where 2 x 136 KB does not
    // fit in shared memory: there the stages run in place with one extra barrier each.
    static constexpr bool PINGPONG = (LOG2M <= 13);
    static constexpr int NBUF = PINGPONG ? 2 : 1;
    static_assert(LOG2M >= 8 && LOG2M <= 14, "supported complex sizes: 256 .. 16384");
};

// Per-thread stage twiddles (registers, constant over all frames of a CTA).
template <int LOG2M>
struct StageTw {
    Tw4 t[FftShape<LOG2M>::NTW > 0 ? FftShape<LOG2M>::NTW : 1];
};

// twM: W_M^e = exp(-2 pi i e / M), e in [0, M)
template <int LOG2M>
__device__ __forceinline__ void load_stage_twiddles(StageTw<LOG2M>& st, const float2* __restrict__ twM, int t) {
    using S = FftShape<LOG2M>;
    int s = 1;
#pragma unroll
    for (int i = 0; i < S::NTW; ++i) {
        int e = (t / s) * s;                       // p * s
        st.t[i].w1 = __ldg(twM + e);
        st.t[i].w2 = __ldg(twM + 2 * e);
        st.t[i].w4 = __ldg(twM + 4 * e);
        st.t[i].w8 = __ldg(twM + 8 * e);
        s *= 16;
    }
}

// ---------------------------------------------------------------------------------------------
// The transform proper.  On entry v[16] holds the stage-1 inputs z[t + j*TPF]; buf0/buf1 are the
// calling sub-FFT's two padded buffers.  All threads of the CTA must call this together (it
// contains __syncthreads()).  Returns the buffer that holds Z[0..M) in natural (padded) order;
// the data is visible to all threads on return.
// If KEEP_LAST_IN_REGS the final stage's outputs are NOT written to shared memory but left in
// v[] (used by the true-peak kernel, which only needs max |.| of the last stage's outputs);
// for the tail radix R3 < 16 the outputs of butterfly i are v[i*R3 .. i*R3+R3).
// ---------------------------------------------------------------------------------------------
template <int LOG2M, bool KEEP_LAST_IN_REGS>
__device__ __forceinline__ float2* fft_forward(float2* v, float2* buf0, float2* buf1,
                                               const StageTw<LOG2M>& st, int t, bool active) {
    using S = FftShape<LOG2M>;
    constexpr int TPF = S::TPF;
    constexpr int RSTRIDE = TPF + TPF / 16;       // padded distance between a thread's 16 inputs
    float2* src = buf0;
    float2* dst = buf0;
    int s = 1;
    const int rbase = padi(t);
#pragma unroll
    for (int i = 0; i < S::NTW; ++i) {
        if (i > 0) {
            if (active) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = src[rbase + j * RSTRIDE];
            }
            if (!S::PINGPONG) __syncthreads();     // in place: everyone has read before anyone writes
        }
        dst = (S::PINGPONG && (i & 1)) ? buf1 : buf0;
        if (active) {
            bf16pt(v);
            apply_twiddles(v, st.t[i]);
            int q = t % s, p = t / s;
            if (s == 1) {
                int wb = 17 * t;
#pragma unroll
                for (int k = 0; k < 16; ++k) dst[wb + k] = v[k];
            } else {
                int wb = padi(q + 16 * s * p);
                int ws = s + s / 16;
#pragma unroll
                for (int k = 0; k < 16; ++k) dst[wb + k * ws] = v[k];
            }
        }
        __syncthreads();
        src = dst;
        s *= 16;
    }
    // remaining sub-length n = M / s is 16 (LAST16), or TAIL in {2,4,8}, or both are absent
    if (S::LAST16) {
        if (active) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = src[rbase + j * RSTRIDE];
            bf16pt(v);
            if (!KEEP_LAST_IN_REGS) {
#pragma unroll
                for (int k = 0; k < 16; ++k) src[rbase + k * RSTRIDE] = v[k];   // same slots: in place
            }
        }
    } else {
        constexpr int R3 = S::TAIL;
        constexpr int SS = S::M / R3;                 // stride between a tail butterfly's points
        constexpr int SSP = SS + SS / 16;
        if (active) {
#pragma unroll
            for (int i = 0; i < 16 / R3; ++i) {
                int qb = padi(t + TPF * i);
#pragma unroll
                for (int j = 0; j < R3; ++j) v[i * R3 + j] = src[qb + j * SSP];
                if (R3 == 2) bf2(v[i * 2], v[i * 2 + 1]);
                if (R3 == 4) bf4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
                if (R3 == 8) bf8(v + i * 8);
                if (!KEEP_LAST_IN_REGS) {
#pragma unroll
                    for (int j = 0; j < R3; ++j) src[qb + j * SSP] = v[i * R3 + j];
                }
            }
        }
    }
    if (!KEEP_LAST_IN_REGS) __syncthreads();
    return src;
}

// ---------------------------------------------------------------------------------------------
// "Local" variant for M <= 4096 (LOG2M 8..12): only ONE all-to-all exchange.
// After the first radix-16 DIF stage the transform splits into 16 independent sub-FFTs of
// n2 = M/16 points (output index = q mod 16).  Sub-FFT q is owned by G2 = M/256 CONSECUTIVE
// threads (<= 16, inside one warp), so its radix-16 stage, the 16 x G2 exchange and the final
// radix-G2 stage need only __syncwarp().  CTA barriers per transform: 2 (after the stage-1 store,
// after the final store) instead of 4 -- ncu showed 21 % of the stall samples at barriers.
//   X buffer (BUF float2, "grouped"): sub-FFT q lives at X[q*S .. q*S+n2), S = 17*G2; the same
//       region is reused for the exchange with slot(p,k) = 17p + k.  All accesses conflict free.
//   Z buffer (M float2, natural order, XOR swizzle zaddr()): conflict free for the final store
//       (lanes vary in q and p) and for consecutive reads by the epilogue.
// ---------------------------------------------------------------------------------------------
// Synchronise the NTHREADS consecutive threads that own one sub-FFT ("group" index g inside the
// CTA): a warp-level sync when the group fits in a warp, a named barrier when it spans 2..4
// warps, the CTA barrier when the group is the whole 256-thread CTA.  Groups of different frames
// never wait for each other.
template <int NTHREADS>
__device__ __forceinline__ void group_sync(int g) {
    if (NTHREADS <= 32) {
        __syncwarp();
    } else if (NTHREADS < 256) {
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(NTHREADS) : "memory");
    } else {
        __syncthreads();
    }
}

template <int LOG2M>
__host__ __device__ __forceinline__ constexpr int zaddr(int idx) {
    constexpr int G2 = (1 << LOG2M) / 256;
    return (idx & ~15) | ((idx & 15) ^ ((((idx >> 4) & (G2 - 1)) * (16 / G2)) & 15));
}

template <int LOG2M>
struct LocalTw { Tw4 s1, s2; };
// same products as LocalTw (bit-identical results), expanded once per CTA
template <int LOG2M>
struct LocalTwFull { Tw15 s1, s2; };
template <int LOG2M>
__device__ __forceinline__ void expand_local_twiddles(const LocalTw<LOG2M>& a, LocalTwFull<LOG2M>& o) {
    expand_twiddles(a.s1, o.s1);
    expand_twiddles(a.s2, o.s2);
}

template <int LOG2M>
__device__ __forceinline__ void load_local_twiddles(LocalTw<LOG2M>& st, const float2* __restrict__ twM, int t) {
    constexpr int G2 = (1 << LOG2M) / 256;
    st.s1.w1 = __ldg(twM + t);     st.s1.w2 = __ldg(twM + 2 * t);
    st.s1.w4 = __ldg(twM + 4 * t); st.s1.w8 = __ldg(twM + 8 * t);
    const int e = 16 * (t % G2);
    st.s2.w1 = __ldg(twM + e);     st.s2.w2 = __ldg(twM + 2 * e);
    st.s2.w4 = __ldg(twM + 4 * e); st.s2.w8 = __ldg(twM + 8 * e);
}

// On entry v[16] = z[t + j*M/16].  X, Z: this sub-FFT group's buffers.  Contains two group
// barriers (after the stage-1 store; after the final store unless KEEP_LAST_IN_REGS); all
// threads of the group must call it together, groups are independent of each other.
// MID is invoked by every thread right after the first barrier (used to overlap deferred work).
// With KEEP_LAST_IN_REGS the final outputs stay in v[] (butterfly i of the last stage in
// v[i*G2 .. i*G2+G2)) and nothing is written to Z.
// zmask (group-uniform): bit pq set = the outputs k in [256 pq, 256 pq + 256) are stored to Z; a caller that reads only
// part of the spectrum (fused epilogue) skips the stores of the other blocks.
#ifndef FFT_LAST_STAGE_SHFL
#define FFT_LAST_STAGE_SHFL 1
#endif
template <int LOG2M, bool KEEP_LAST_IN_REGS, typename TW, typename Mid>
__device__ __forceinline__ void fft_forward_local(float2* v, float2* X, float2* Z, const TW& st,
                                                  int t, int g, bool active, Mid&& mid, unsigned zmask = 0xffffffffu) {
    constexpr int M = 1 << LOG2M;
    constexpr int TPF = M / 16;
    constexpr int G2 = M / 256;
    constexpr int S = 17 * G2;
    static_assert(LOG2M >= 8 && LOG2M <= 12, "local variant covers 256 .. 4096 complex points");
    if (active) {
        bf16pt(v);
        apply_twiddles(v, st.s1);
#pragma unroll
        for (int k = 0; k < 16; ++k) X[k * S + t] = v[k];
    }
    group_sync<TPF>(g);
    mid();
    if (active) {
        const int q = t / G2, p = t % G2;
        float2* Xq = X + q * S;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = Xq[p + G2 * j];
        bf16pt(v);
        if (G2 == 1) {
            if (!KEEP_LAST_IN_REGS) {
#pragma unroll
                for (int k = 0; k < 16; ++k) Z[zaddr<LOG2M>(q + 16 * k)] = v[k];
            }
        } else if (G2 == 2 && FFT_LAST_STAGE_SHFL) {
            // N = 1024: the last stage is a radix-2 butterfly between the two lanes of a pair (p = t & 1).  Each lane
            // keeps the half of its values whose index has its own parity and trades the other half with its
            // neighbour through the register crossbar (16 SHFL) instead of shared memory (16 STS.64 + 16 LDS.64):
            // out0 = a + b is symmetric in the two lanes' values, out1 = a - b changes sign with the lane's parity --
            // the same sums as the shared-memory exchange, bit for bit.
            apply_twiddles(v, st.s2);
            const int zq = zaddr<LOG2M>(q + 16 * p);
            const float sgn = p ? -1.f : 1.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 keep = p ? v[2 * i + 1] : v[2 * i];          // v_p[p + 2 i]
                const float2 send = p ? v[2 * i] : v[2 * i + 1];          // v_p[(1 - p) + 2 i]: the neighbour's index
                float2 recv;
                recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                v[2 * i] = cadd(keep, recv);                               // a + b
                v[2 * i + 1] = cmul_elem(csub(keep, recv), make_float2(sgn, sgn));      // a - b = +-(keep - recv)
                if (!KEEP_LAST_IN_REGS) {
#pragma unroll
                    for (int pq = 0; pq < 2; ++pq)
                        if (zmask & (1u << pq)) Z[zq + 32 * i + 256 * pq] = v[2 * i + pq];
                }
            }
        } else {
            apply_twiddles(v, st.s2);
            __syncwarp();                         // every lane of the group holds its inputs in registers
#pragma unroll
            for (int k = 0; k < 16; ++k) Xq[17 * p + k] = v[k];
            __syncwarp();
            // zaddr(q + 16 (p + G2 i) + 256 pq) = zq + 16 G2 i + 256 pq: the swizzle term depends on p and q only
            // (q < 16 fills bits 0-3, p the next log2 G2 bits), so the bit arithmetic is done once per thread
            const int zq = zaddr<LOG2M>(q + 16 * p);
#pragma unroll
            for (int i = 0; i < 16 / G2; ++i) {
                const int k = p + G2 * i;
#pragma unroll
                for (int pp = 0; pp < G2; ++pp) v[i * G2 + pp] = Xq[17 * pp + k];
                if (G2 == 2) bf2(v[i * 2], v[i * 2 + 1]);
                if (G2 == 4) bf4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
                if (G2 == 8) bf8(v + i * 8);
                if (G2 == 16) bf16pt(v);
                if (!KEEP_LAST_IN_REGS) {
#pragma unroll
                    for (int pq = 0; pq < G2; ++pq)
                        if (zmask & (1u << pq)) Z[zq + 16 * G2 * i + 256 * pq] = v[i * G2 + pq];
                }
            }
        }
    }
    if (!KEEP_LAST_IN_REGS) group_sync<TPF>(g);
}

// Untangle one (k, M-k) pair of the packed real transform.
//   Zk = Z[k], Zm = Z[(M-k) % M], w = W_N^k = exp(-2 pi i k / N), N = 2M
//   X[k] = E + w*O,  X[M-k] = conj(E - w*O),  E = (Zk + conj Zm)/2,  O = -i (Zk - conj Zm)/2
__device__ __forceinline__ void rfft_pair(float2 Zk, float2 Zm, float2 w, float2& Xk, float2& Xmk) {
    // E = ((Zk.x + Zm.x) / 2, (Zk.y - Zm.y) / 2), O = ((Zk.y + Zm.y) / 2, -(Zk.x - Zm.x) / 2) as packed operations
    const float2 hf = make_float2(0.5f, 0.5f);
    float2 E = __fmul2_rn(__fadd2_rn(make_float2(Zm.x, -Zm.y), Zk), hf);
    float2 O = __fmul2_rn(__fadd2_rn(make_float2(Zk.y, -Zk.x), make_float2(Zm.y, Zm.x)), hf);
    float2 T = cmul(w, O);
    Xk = cadd(E, T);
    float2 D = csub(E, T);
    Xmk = make_float2(D.x, -D.y);
}

__device__ __forceinline__ float cabs(float2 a) { return sqrtf(fmaf(a.x, a.x, a.y * a.y)); }
// |a| with the hardware square root (one MUFU.SQRT, relative error 2^-23 against 1.2e-3 for the 0.01 dB bar;
// the IEEE sqrtf above costs ~9 instructions and a divergent slow path per call); squares below 1.2e-38 flush to 0
__device__ __forceinline__ float cabs_fast(float2 a) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaf(a.x, a.x, a.y * a.y)));
    return y;
}

}  // namespace o4
