// blockdft_tc_kernel.cuh -- the hop-block partial DFT GEMM of blockdft_kernel.cuh on the 5th-gen
// tensor cores: tcgen05.mma kind::tf32 with fp32 accumulators in TMEM, 3xTF32 split precision.
//
//   Q[blocks x cols] = X[blocks x hop] . E[hop x cols]            (cols <= 2048: up to eight 256-column tiles)
//
// TF32 keeps 11 significant bits, far too few for the 0.01 dB parity bar next to strong peaks, so
// both operands are split exactly,  x = x_hi + x_lo,  e = e_hi + e_lo  (hi = the top 11 bits, lo = the
// remainder, itself rounded to TF32 by the hardware), and three products are accumulated in fp32:
//   x_hi e_hi + x_hi e_lo + x_lo e_hi        (the dropped x_lo e_lo term is 2^-22 relative)
// which restores ~fp32 accuracy at 3x the tensor work -- still several times cheaper than the CUDA-core
// GEMM or the FFT for the resolutions that need <= 1024 columns.  The tensor-core accumulator truncates
// instead of rounding (measured: error grows linearly with the number of accumulation steps), so the
// large x_hi e_hi products and the 2^-11 smaller cross products go to SEPARATE TMEM accumulators and
// are added once, in registers, in the epilogue: the main accumulator then sees 64 steps, not 192.
//
// One CTA = 128 hop blocks x 256 columns (TC_MH = 1; or 256 x 128 with TC_MH = 2); TMEM: [main | cross]
// x 256 columns = all 512.  K in chunks of 16 samples.  Shared-memory operand tiles use the canonical K-major SWIZZLE_64B layout
// (8-row groups of 64-byte rows, 16-byte chunk index XOR (row >> 1) & 3, SBO = 512 B):
//   A_hi / A_lo: written by four producer groups of two warps, one chunk in flight per group
//                (LDG.128 of the raw samples, split, STS.128, fence.proxy.async, one arrive per warp)
//   B_hi / B_lo: the constant E table, pre-swizzled on the host into per-chunk byte images and
//                fetched with one-dimensional cp.async.bulk (TMA without a tensor map) + mbarrier tx
// Warp 8's lane 0 issues the MMAs (6 per chunk) and commits them to the stage's mbarrier; warp 9's lane 0
// requests the B images as stages retire (no block-wide barrier in the main loop).
// Epilogue: tcgen05.ld 32x32b -> registers (main + cross), then either Q as it is, or (exact-windowing operand)
// the frame sums X_f[k] = sum_b Q_{f+1-B+b}[k][b] through shared memory -> X[ch][bin][frame].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace o4 {

#ifndef TC_MH
#define TC_MH 1                     // row halves of 128 hop blocks per CTA (1 or 2); TC_MH * TC_BN = 256
#endif
constexpr int TC_BM = 128 * TC_MH;  // hop blocks per CTA
constexpr int TC_BN = 256 / TC_MH;  // columns per CTA (UMMA N)
constexpr int TC_KC = 16;           // samples per K chunk: 64-byte rows
constexpr int TC_STAGES = 4;
constexpr int TC_MAX_GROUPS = 64;     // 32-column groups of the fused epilogue's tables: up to 2048 GEMM columns
#ifndef TC_EXPERIMENT_KS
#define TC_EXPERIMENT_KS 2          // timing experiments only: 1 issues half of the MMAs (wrong results)
#endif
// producer groups = chunks produced side by side.  Must equal TC_STAGES: a group then owns one stage and sees
// every phase of its mbarriers; with any other ratio a group visits a stage only now and then and the parity
// wait can mistake an older phase for the one it needs (tried: 4 groups x 3 stages dead-locks).
#define TC_PG TC_STAGES
constexpr int TC_THREADS = 256;
constexpr int TC_A_BYTES = TC_BM * TC_KC * 4;     // 8 / 16 KB per (hi | lo)
constexpr int TC_B_BYTES = TC_BN * TC_KC * 4;     // 16 / 8 KB per (hi | lo)
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 48 KB

struct BlockDftTcArgs {
    const float* x;            // samples; block j of channel c = x + c*ch_stride + j*hop
    long long ch_stride;
    int hop;                   // K, multiple of TC_KC
    int n_ch;
    int j0, nb;
    int n_halves;              // column tiles of TC_BN (1 .. 8)
    const uint8_t* Eimg;       // [n_halves][hop/KC][2 (hi, lo)][TC_B_BYTES] pre-swizzled operand images
    float* Q;                  // [n_ch][nb][qs]
    int qs;                    // TC_BN * n_halves
    // Fused frame assembly (exact-windowing operand only, see blockdft_kernel.cuh): instead of writing Q, the
    // epilogue adds the block rows of every frame, X_f[k] = sum_b Q_{f+1-B+b}[k][b], through shared memory
    // and writes the complex bins X[ch][bin][frame]; Q never leaves the SM.  Frames that straddle two row
    // tiles are completed with atomicAdd on the zero-initialised X (B - 1 of 128 frames per tile and side).
    float2* X;                 // nullptr: plain Q output
    int n_frames;              // frames (= hops) per channel; frame f ends with hop block f
    int nkx;                   // bins per frame over all fused resolutions
    short gB[TC_MAX_GROUPS];   // per 32-column group (n_halves * 8): block positions per bin (2, 4, 8, 16; 0 = unused group)
    short gX[TC_MAX_GROUPS];   // per group: index of its first bin in [0, nkx)
    short gN[TC_MAX_GROUPS];   // per group: bins present (<= 16 / B)
    short gS[TC_MAX_GROUPS];   // per group: frame shift.  A transform of more than 16 hop blocks (N / hop = 32, 64) spreads a
                               // bin over N / hop / 16 groups of 16 positions; the group holding positions 16 p .. 16 p + 15
                               // adds its partial sum to the frame that ends gS = N / hop - 16 (p + 1) blocks after the
                               // group's own last row, always with atomicAdd (several groups feed one X element); -1 = plain group
};
constexpr int TC_XROW = 17;          // float2 row stride of the epilogue exchange strip (conflict free)

// byte offset of (row, 16-byte chunk c in 0..3) inside a K-major SWIZZLE_64B tile of 64-byte rows
__host__ __device__ __forceinline__ int tc_sw64_offset(int row, int c) {
    return (row >> 3) * 512 + (row & 7) * 64 + ((c ^ ((row >> 1) & 3)) << 4);
}

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = tc_smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(a), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc_smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc_smem_u32(bar)) : "memory");
}

// K-major SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t tc_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);            // start address, 16-byte units
    d |= (uint64_t)1 << 16;                                // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;                       // stride byte offset: 8-row group pitch
    d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;                                // layout type SWIZZLE_64B
    return d;
}

// instruction descriptor: D = F32, A = B = TF32, K-major both, N = TC_BN, M = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tc_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}


// Fused frame assembly of one 32-column group: S holds the 128 rows x 16 complex columns (bin q, block position
// b at column q B + b).  Task 0: the frame whose last hop block is local row r; task 1 (r < B - 1): the frame that
// ends r + 1 rows past the tile.  Frames that lie wholly inside the tile are stored, the others added atomically.
template <int B>
__device__ __forceinline__ void tc_frame_sums(const BlockDftTcArgs& a, const float2* S, int r, int row0, int ch, int nbin, int xb0,
                                              int shift = -1) {
    constexpr int PER = 16 / B;
#pragma unroll
    for (int task = 0; task < 2; ++task) {
        if (task == 1 && r >= B - 1) break;
        const int R = task == 0 ? r : 128 + r;
        const long long f = (long long)a.j0 + row0 + R + (shift > 0 ? shift : 0);   // frame index = index of its last hop block
        if (f < 0 || f >= a.n_frames) continue;
        const bool whole = (R - B + 1 >= 0) && (R <= 127) && shift < 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            if (q >= nbin) break;
            float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const int rho = R - B + 1 + b;
                float2 u = make_float2(0.f, 0.f);
                if (rho >= 0 && rho <= 127) u = S[rho * TC_XROW + q * B + b];
                if (b & 1) { acc1.x += u.x; acc1.y += u.y; } else { acc0.x += u.x; acc0.y += u.y; }
            }
            const float2 acc = make_float2(acc0.x + acc1.x, acc0.y + acc1.y);
            float2* dst = a.X + ((size_t)ch * a.nkx + xb0 + q) * a.n_frames + f;
            if (whole) *dst = acc;
            else { atomicAdd(&dst->x, acc.x); atomicAdd(&dst->y, acc.y); }
        }
    }
}

// Warp roles: warps 0..7 (256 threads) produce the A tiles and later run the epilogue; warp 8, lane 0 issues
// every tcgen05.mma and waits for operands only; warp 9, lane 0 requests the B images (bulk TMA) as stages
// retire (one thread doing both had to wait for chunk k-2 to COMPLETE before issuing chunk k, which left
// the tensor pipe one chunk of work to hide the completion round trip).  Three mbarrier rings tie them together:
//   bar_a[s]  8 / TC_PG arrivals (one per warp of the producer group) A_hi / A_lo of the stage are written and fenced
//   bar_b[s]  tx bytes      the B image of the stage has landed
//   bar_m[s]  tcgen05.commit: every MMA issued so far (in particular those reading stage s) has retired
__global__ void __launch_bounds__(TC_THREADS + 64, 1)
blockdft_tc_kernel(const __grid_constant__ BlockDftTcArgs a) {
    extern __shared__ __align__(16) uint8_t tc_smem_raw[];
    // swizzled operand tiles need a 1024-byte aligned base (the swizzle is a function of address bits)
    // (offset added to the __shared__ array itself, not a round trip through uintptr_t: the compiler then keeps the
    // shared address space and emits STS.128 / LDS.64 instead of four generic 32-bit stores per 16-byte piece --
    // ncu r02a: 126 M of the kernel's 238 M shared wavefronts were that 4-way split)
    uint8_t* tiles = tc_smem_raw + ((1024u - (tc_smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    uint64_t* bar_a = reinterpret_cast<uint64_t*>(tiles + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* bar_b = bar_a + TC_STAGES;
    uint64_t* bar_m = bar_b + TC_STAGES;
    uint64_t* bar_done = bar_m + TC_STAGES;       // one commit after the last MMA: the epilogue's start signal
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_per_ch = (a.nb + TC_BM - 1) / TC_BM;
    int bid = blockIdx.x;
    const int half = bid % a.n_halves; bid /= a.n_halves;
    const int row0 = (bid % tiles_per_ch) * TC_BM;
    const int ch = bid / tiles_per_ch;
    const float* xa = a.x + (long long)ch * a.ch_stride + (long long)a.j0 * a.hop;
    const int n_kc = a.hop / TC_KC;
    const uint8_t* eimg = a.Eimg + (size_t)half * n_kc * 2 * TC_B_BYTES;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            tc_mbar_init(bar_a + s, TC_THREADS / 32 / TC_PG); tc_mbar_init(bar_b + s, 1); tc_mbar_init(bar_m + s, 1);
        }
        tc_mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc_smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == TC_THREADS / 32 + 1) {
        // ===== B-operand issuer (one thread): runs up to STAGES chunks ahead of the MMAs' completion =====
        if (lane == 0) {
            for (int kb = 0; kb < n_kc; ++kb) {
                const int sb = kb % TC_STAGES;
                // the stage was last read by chunk kb - STAGES
                if (kb >= TC_STAGES) tc_mbar_wait(bar_m + sb, (uint32_t)(kb / TC_STAGES - 1) & 1);
#ifdef TC_EXPERIMENT_B_DIV      /* timing experiment only (wrong results): fetch 1/DIV of the B bytes */
                tc_mbar_expect_tx(bar_b + sb, 2 * TC_B_BYTES / TC_EXPERIMENT_B_DIV);
                tc_bulk_g2s(tiles + sb * TC_STAGE_BYTES + 2 * TC_A_BYTES, eimg + (size_t)kb * 2 * TC_B_BYTES, 2 * TC_B_BYTES / TC_EXPERIMENT_B_DIV, bar_b + sb);
#else
                tc_mbar_expect_tx(bar_b + sb, 2 * TC_B_BYTES);
                tc_bulk_g2s(tiles + sb * TC_STAGE_BYTES + 2 * TC_A_BYTES, eimg + (size_t)kb * 2 * TC_B_BYTES, 2 * TC_B_BYTES, bar_b + sb);
#endif
            }
        }
    } else if (warp == TC_THREADS / 32) {
        // ===== MMA issuer (one thread): never waits for a completion, only for operands =====
        if (lane == 0) {
            for (int kc = 0; kc < n_kc; ++kc) {
                const int s = kc % TC_STAGES;
                const uint32_t use = (uint32_t)(kc / TC_STAGES);
                tc_mbar_wait(bar_a + s, use & 1);
                tc_mbar_wait(bar_b + s, use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = tc_smem_u32(tiles + s * TC_STAGE_BYTES);
#pragma unroll
                for (int mh = 0; mh < TC_MH; ++mh) {
#pragma unroll
                    for (int ks = 0; ks < TC_EXPERIMENT_KS; ++ks) {
                        const uint64_t d_ahi = tc_desc_sw64(sa + mh * (TC_A_BYTES / TC_MH) + ks * 32);
                        const uint64_t d_alo = tc_desc_sw64(sa + TC_A_BYTES + mh * (TC_A_BYTES / TC_MH) + ks * 32);
                        const uint64_t d_bhi = tc_desc_sw64(sa + 2 * TC_A_BYTES + ks * 32);
                        const uint64_t d_blo = tc_desc_sw64(sa + 2 * TC_A_BYTES + TC_B_BYTES + ks * 32);
                        const uint32_t t_main = tmem + mh * TC_BN, t_cross = tmem + 256 + mh * TC_BN;
                        tc_mma_tf32(t_main, d_ahi, d_bhi, (kc | ks) != 0);
                        tc_mma_tf32(t_cross, d_ahi, d_blo, (kc | ks) != 0);
                        tc_mma_tf32(t_cross, d_alo, d_bhi, 1);
                    }
                }
                tc_commit(bar_m + s);
            }
            tc_commit(bar_done);
        }
    } else {
        // ===== A producers.  TC_PG groups of 256 / TC_PG threads each own every TC_PG-th chunk: a chunk costs its
        // producer a wait + STS + fence.proxy.async + arrive round trip of ~1000 cycles that no amount of
        // look-ahead shortens (measured: halving the MMAs or the B bytes changed nothing, the chain did), so
        // TC_PG chunks are kept in flight side by side instead of one after the other =====
        constexpr int PGT = TC_THREADS / TC_PG;                  // threads per producer group
        constexpr int NQ = TC_BM * 4 / PGT;                      // 16-byte pieces per thread and chunk
        const int pg = tid / PGT, tg = tid % PGT;
        const float4* a_src[NQ];
        int a_off[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int p = tg + PGT * q;
            const int row = p >> 2, c = p & 3;
            int r = row0 + row;
            if (r >= a.nb) r = a.nb - 1;                      // rows past the end: computed, never stored
            a_src[q] = reinterpret_cast<const float4*>(xa + (long long)r * a.hop) + c;
            a_off[q] = tc_sw64_offset(row, c);
        }
        float4 nxt[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) nxt[q] = (pg < n_kc) ? __ldg(a_src[q] + pg * (TC_KC / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kc = pg; kc < n_kc; kc += TC_PG) {
            const int s = kc % TC_STAGES;
            uint8_t* st = tiles + s * TC_STAGE_BYTES;
            float4 cur[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) cur[q] = nxt[q];
#ifndef TC_EXPERIMENT_NO_LDG
            if (kc + TC_PG < n_kc) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) nxt[q] = __ldg(a_src[q] + (kc + TC_PG) * (TC_KC / 4));
            }
#endif
            // the MMAs that read this stage (chunk kc - STAGES) must have retired before it is overwritten
            if (kc >= TC_STAGES) tc_mbar_wait(bar_m + s, (uint32_t)(kc / TC_STAGES - 1) & 1);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(cur[q].x) & 0xFFFFE000u); lo.x = cur[q].x - hi.x;
                hi.y = __uint_as_float(__float_as_uint(cur[q].y) & 0xFFFFE000u); lo.y = cur[q].y - hi.y;
                hi.z = __uint_as_float(__float_as_uint(cur[q].z) & 0xFFFFE000u); lo.z = cur[q].z - hi.z;
                hi.w = __uint_as_float(__float_as_uint(cur[q].w) & 0xFFFFE000u); lo.w = cur[q].w - hi.w;
                *reinterpret_cast<float4*>(st + a_off[q]) = hi;
                *reinterpret_cast<float4*>(st + TC_A_BYTES + a_off[q]) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> tensor-core reads
            __syncwarp();                                                      // one arrival per warp
            if (lane == 0) tc_mbar_arrive(bar_a + s);
        }
        // all MMAs retired <=> the final commit has arrived.  (Not bar_m of the last stage: a producer group that
        // never waited on that stage inside the loop could see an older phase of the same parity as complete.)
        tc_mbar_wait(bar_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31; the two warp groups split the row halves
        // (TC_MH = 2) or the columns (TC_MH = 1); main + cross accumulators are added in registers
        const int grp = warp >> 2, quad = warp & 3;
        const int mh = (TC_MH == 2) ? grp : 0;
        const int c0 = (TC_MH == 2) ? 0 : grp * (TC_BN / 2);
        constexpr int NCG = (TC_MH == 2) ? TC_BN / 32 : TC_BN / 64;
        const int row = row0 + mh * 128 + quad * 32 + lane;
        if (TC_MH == 1 && a.X) {
            // ---- fused frame assembly.  The 128 threads of a group own the 128 rows of the tile for their
            // 128 columns; the operand stages are idle now and serve as the exchange strip.
            float2* S = reinterpret_cast<float2*>(tiles) + grp * (128 * TC_XROW);
            const int r = quad * 32 + lane;                           // local row = local index of the frame's last block
#pragma unroll 1
            for (int cg = 0; cg < NCG; ++cg) {
                uint32_t v[32], c[32];
                const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c0 + cg * 32);
                tc_tmem_ld32(taddr, v);
                tc_tmem_ld32(taddr + 256, c);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int gi = half * (TC_BN / 32) + (c0 >> 5) + cg;
                const int B = a.gB[gi], nbin = a.gN[gi], xb0 = a.gX[gi];
#ifdef TC_EXPERIMENT_NO_EPI
                if (B > 1000) {
#else
                if (B > 0) {
#endif
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        S[r * TC_XROW + i] = make_float2(__uint_as_float(v[2 * i]) + __uint_as_float(c[2 * i]),
                                                         __uint_as_float(v[2 * i + 1]) + __uint_as_float(c[2 * i + 1]));
                }
                asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
#ifdef TC_EXPERIMENT_NO_EPI
                if (B > 1000) {
#else
                if (B > 0) {
#endif
                    if (B == 16) tc_frame_sums<16>(a, S, r, row0, ch, nbin, xb0, a.gS[gi]);
                    else if (B == 8) tc_frame_sums<8>(a, S, r, row0, ch, nbin, xb0);
                    else if (B == 4) tc_frame_sums<4>(a, S, r, row0, ch, nbin, xb0);
                    else tc_frame_sums<2>(a, S, r, row0, ch, nbin, xb0);
                }
                asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
            }
        } else {
            float* qrow = a.Q + ((size_t)ch * a.nb + (row < a.nb ? row : 0)) * a.qs + half * TC_BN + c0;
#pragma unroll 1
            for (int cg = 0; cg < NCG; ++cg) {
                uint32_t v[32], c[32];
                const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mh * TC_BN + c0 + cg * 32);
                tc_tmem_ld32(taddr, v);
                tc_tmem_ld32(taddr + 256, c);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < a.nb) {
                    float4* dst = reinterpret_cast<float4*>(qrow + cg * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        dst[i] = make_float4(__uint_as_float(v[4 * i]) + __uint_as_float(c[4 * i]),
                                             __uint_as_float(v[4 * i + 1]) + __uint_as_float(c[4 * i + 1]),
                                             __uint_as_float(v[4 * i + 2]) + __uint_as_float(c[4 * i + 2]),
                                             __uint_as_float(v[4 * i + 3]) + __uint_as_float(c[4 * i + 3]));
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

inline size_t blockdft_tc_smem_bytes() { return (size_t)TC_STAGES * TC_STAGE_BYTES + (3 * TC_STAGES + 1) * 8 + 16 + 1024; }

}  // namespace o4
