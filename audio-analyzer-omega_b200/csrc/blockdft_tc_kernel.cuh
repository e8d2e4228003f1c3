// blockdft_tc_kernel.cuh -- the hop-block partial DFT GEMM of blockdft_kernel.cuh on the 5th-gen
// tensor cores: tcgen05.mma with fp32 accumulators in TMEM, split-precision operands.
//
//   Q[blocks x cols] = X[blocks x hop] . E[hop x cols]            (cols <= 2048: up to eight 256-column tiles)
//
// A tensor-core input keeps 11 significant bits, far too few for the 0.01 dB parity bar next to strong peaks, so
// both operands are split,  x = x_hi + x_lo,  e = e_hi + e_lo, and three products are accumulated in fp32:
//   x_hi e_hi + x_hi e_lo + x_lo e_hi        (the dropped x_lo e_lo term is 2^-22 relative)
// which restores ~fp32 accuracy at 3x the tensor work -- still several times cheaper than the CUDA-core
// GEMM or the FFT for the resolutions that need <= 1024 columns.  The tensor-core accumulator truncates
// instead of rounding (measured: error grows linearly with the number of accumulation steps), so the
// large x_hi e_hi products and the 2^-11 smaller cross products go to SEPARATE TMEM accumulators and
// are added once, in registers, in the epilogue.
//
// Operand precision (both reach ~2^-22 per element, i.e. float32-grade products accumulated in fp32):
//   TC_F16 0: kind::tf32, hi = the top 11 bits, lo = the remainder (3xTF32), 16 samples per 64-byte operand row
//   TC_F16 1: kind::f16,  x s = hi + lo / 2^11 in IEEE half (11 + 11 significant bits), 32 samples per 64-byte row:
//             the same bytes, barriers and MMAs per chunk carry twice the samples, and an f16 MMA takes as long as
//             a tf32 one -- half the tensor time per sample (the K loop runs at the MMA rate: 768 cycles per
//             chunk, profiles/r02j_tc_timeline.txt).  Half has 5 exponent bits, so every hop-block row is first
//             scaled by a power of two to a maximum in [0.5, 1) (hopblock_scale_kernel; exact, undone in the
//             epilogue), the low parts are stored times 2^11 and their accumulator is scaled back in the epilogue.
//             Elements more than 2^14 below their row's maximum go subnormal in hi and are picked up by lo to
//             2^-36 of the row maximum, so the result is scale invariant like the float32 reference.
//
// PERSISTENT, warp-specialised: one CTA per SM walks over (row tile, column tile) pairs of 128 hop blocks x 256
// columns; TMEM: [main | cross] x 256 columns = all 512, allocated once.  K in chunks of 64-byte operand rows.
// Shared-memory operand tiles use the canonical K-major SWIZZLE_64B layout (8-row groups of 64-byte rows, 16-byte
// chunk index XOR (row >> 1) & 3, SBO = 512 B), three stages of 48 KB:
//   warps 0-5   A producers, three groups of two warps, one chunk in flight per group: LDG.128 of the raw samples,
//               scale + split, STS.128, fence.proxy.async, one mbarrier arrive per warp
//   warp 6      lane 0 issues the MMAs (6 per chunk) and commits them to the stage's mbarrier
//   warp 7      lane 0 requests the B images (the constant E table, pre-swizzled on the host into per-chunk byte
//               images; one-dimensional cp.async.bulk + mbarrier tx) as stages retire
//   warps 8-15  epilogue, two groups of four warps (128 of the 256 columns each, own exchange strip): tcgen05.ld
//               32x32b -> registers (main + cross), then either Q as it is, or (exact-windowing operand) the frame
//               sums X_f[k] = sum_b Q_{f+1-B+b}[k][b] through the strip -> X[ch][bin][frame].  While they drain a
//               tile, the producers and the B issuer already stage the next tile's first chunks; the MMA issuer
//               waits for the TMEM to be free (bar_empty) and goes on.
// Measured with clock stamps (profiles/r02j_tc_timeline.txt): when every tile was its own CTA, TMEM allocation,
// barrier set-up and first operand fetch (5.9 k cycles), an epilogue by the producer warps (7.1 k, most of it
// load -> add chains behind predicated shared loads) and the launch gap (2.9 k) were serial with the 12.8 k cycles of
// MMAs (half operands; 24.6 k with 3xTF32).  Now a tile takes 20.7 k: 13.6 k of MMAs + 0.4 k commit + 6.3 k until the
// accumulators are out of TMEM (64 B/clk of TMEM reads bound that at ~4 k).  Two CTAs per SM on 128-column tiles were
// measured slower (7.2 vs 6.3 ms with 3xTF32): an N = 128 MMA reads 8 KB of operands per 64 cycles, all the shared
// memory bandwidth there is.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace o4 {

#ifndef TC_F16
#define TC_F16 1
#endif
constexpr int TC_BM = 128;          // hop blocks per tile (UMMA M)
constexpr int TC_BN = 256;          // columns per tile (UMMA N)
constexpr int TC_KC = TC_F16 ? 32 : 16;   // samples per K chunk: 64-byte rows
constexpr float TC_LO_SCALE = 2048.f;     // TC_F16: low parts are stored times 2^11
constexpr int TC_STAGES = 3;
constexpr int TC_TMEM_COLS = 2 * TC_BN;   // [main | cross]
constexpr int TC_MAX_GROUPS = 64;     // 32-column groups of the fused epilogue's tables: up to 2048 GEMM columns
#ifndef TC_EXPERIMENT_KS
#define TC_EXPERIMENT_KS 2          // timing experiments only: 1 issues half of the MMAs (wrong results)
#endif
// producer groups = chunks produced side by side.  Must equal TC_STAGES: a group then owns one stage and sees
// every phase of its mbarriers; with any other ratio a group visits a stage only now and then and the parity
// wait can mistake an older phase for the one it needs (tried: 4 groups x 3 stages dead-locks).  For the same
// reason a CTA only walks over several tiles when the chunks per tile are a multiple of TC_STAGES (host side).
#define TC_PG TC_STAGES
constexpr int TC_THREADS = 64 * TC_STAGES;         // producer threads: two warps per producer group
constexpr int TC_WARP_MMA = TC_THREADS / 32;       // 6
constexpr int TC_WARP_B = TC_WARP_MMA + 1;         // 7
constexpr int TC_WARP_EPI = TC_WARP_MMA + 2;       // 8 .. 15: warp w reads TMEM lanes 32 (w % 4) .. +31
constexpr int TC_EPI_WARPS = 8;                    // two groups of four warps, each drains 128 of the 256 columns
constexpr int TC_CTA_THREADS = 32 * (TC_WARP_EPI + TC_EPI_WARPS);   // 512
static_assert(TC_WARP_EPI % 4 == 0, "epilogue warp w must own TMEM lane quarter w % 4");
constexpr int TC_A_BYTES = TC_BM * 64;            // 8 KB per (hi | lo)
constexpr int TC_B_BYTES = TC_BN * 64;            // 16 KB per (hi | lo)
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 48 KB
constexpr int TC_XROW = 17;          // float2 row stride of the epilogue exchange strip (conflict free)
constexpr int TC_SPAD = 15;          // all-zero strip rows before and after the tile's 128 (frames of up to 16 hop blocks)
constexpr int TC_STRIP_BYTES = (128 + 2 * TC_SPAD) * TC_XROW * 8;    // per epilogue group

struct BlockDftTcArgs {
    const float* x;            // samples; block j of channel c = x + c*ch_stride + j*hop
    long long ch_stride;
    int hop;                   // K, multiple of TC_KC
    int n_ch;
    int j0, nb;
    int n_halves;              // column tiles of TC_BN (1 .. 8)
    int n_tiles;               // row tiles x channels x column tiles; CTA b takes tiles b, b + gridDim.x, ...
    const uint8_t* Eimg;       // [n_halves][hop/KC][2 (hi, lo)][TC_B_BYTES] pre-swizzled operand images
    const float* row_inv;      // TC_F16: [n_ch][nb] 2^e per hop-block row, the inverse of the row's operand scale
    float e_inv;               // TC_F16: inverse of the power-of-two scale of the E images
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
#ifdef TC_TIMELINE                   /* developer builds: clock stamps of the second tile of every 4th CTA (see omega4_cuda.cu) */
__device__ long long tc_tl[64][16];
#define TC_STAMP(cond, i) do { if ((cond) && (blockIdx.x & 3) == 0 && (blockIdx.x >> 2) < 64) tc_tl[blockIdx.x >> 2][i] = clock64(); } while (0)
#else
#define TC_STAMP(cond, i) do { } while (0)
#endif

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
// (cute::UMMA::InstrDescriptor: c_format [4,6) 1 = F32; a_format [7,10), b_format [10,13): 0 = F16, 2 = TF32)
constexpr uint32_t TC_FMT = TC_F16 ? 0u : 2u;
constexpr uint32_t TC_IDESC = (1u << 4) | (TC_FMT << 7) | (TC_FMT << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
#if TC_F16
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
#else
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
#endif

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

// TC_F16 operand split of two scaled samples: hi = half(a), lo = half((a - hi) 2^11), packed pairs
__device__ __forceinline__ void tc_split_h2(float a0, float a1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a0, a1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn((a0 - hf.x) * TC_LO_SCALE, (a1 - hf.y) * TC_LO_SCALE);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// main + cross accumulator of one output element (TC_F16: the cross sum carries 2^11, the row and the table their scales)
__device__ __forceinline__ float tc_acc(uint32_t v, uint32_t c, float rs) {
#if TC_F16
    return fmaf(__uint_as_float(c), 1.f / TC_LO_SCALE, __uint_as_float(v)) * rs;
#else
    return __uint_as_float(v) + __uint_as_float(c);
#endif
}
// One hop-block row's operand scale for TC_F16: inv = 2^e with max |x| 2^-e in [0.5, 1) (1 for an all-zero row), so that
// half's 5 exponent bits are spent on the 14 binades below the row's own maximum.  One warp per row.
// the scale of a row whose largest magnitude is m
__device__ __forceinline__ float hop_row_scale(float m) {
    int e = (int)((__float_as_uint(m) >> 23) & 0xffu);            // biased exponent: m = [1, 2) 2^(e - 127)
    float inv = 1.f;
    if (m > 0.f && e < 255) {                                     // (Inf / NaN rows keep scale 1: garbage in, garbage out)
        e = e > 250 ? 250 : e;
        inv = __uint_as_float((uint32_t)(e + 1) << 23);           // 2^(e - 126): denormal rows (e = 0) scale by 2^126
    }
    return inv;
}
// inv[ch][0 .. nb) belongs to hop blocks j0 .. j0 + nb - 1; this kernel fills the first `ncols` of them (all of them, or
// only the blocks before the first meter frame when the K-weighting kernel of the same call writes the others: it has
// every later hop block in registers anyway, kweight32_kernel.cuh)
struct HopScaleArgs { const float* x; long long ch_stride; int hop, n_ch, j0, nb, ncols; float* inv; };
__global__ void __launch_bounds__(256)
hopblock_scale_kernel(const HopScaleArgs a) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (long long)a.n_ch * a.ncols) return;
    const int ch = (int)(row / a.ncols), j = (int)(row % a.ncols);
    const float4* px = reinterpret_cast<const float4*>(a.x + (long long)ch * a.ch_stride + ((long long)a.j0 + j) * a.hop);
    float m = 0.f;
    for (int i = lane; i < a.hop / 4; i += 32) {
        const float4 v = __ldg(px + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) a.inv[(size_t)ch * a.nb + j] = hop_row_scale(m);
}


// Fused frame assembly of one 32-column group: S holds the tile's 128 rows x 16 complex columns (bin q, block position
// b at column q B + b) between TC_SPAD all-zero rows on either side, so that the B loads of a frame sum carry no
// predicates (with predicated loads the compiler emitted load -> add -> load chains: ~1000 cycles per group instead
// of ~100).  Task 0: the frame whose last hop block is local row r; task 1 (r < B - 1): the frame that ends r + 1
// rows past the tile.  Frames that lie wholly inside the tile are stored, the others added atomically.
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
        const float2* Sf = S + (R - B + 1 + TC_SPAD) * TC_XROW;                     // row of the frame's first hop block
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            if (q >= nbin) break;
            float2 u[B];
#pragma unroll
            for (int b = 0; b < B; ++b) u[b] = Sf[b * TC_XROW + q * B + b];
#pragma unroll
            for (int w = B / 2; w > 0; w >>= 1)
#pragma unroll
                for (int b = 0; b < w; ++b) { u[b].x += u[b + w].x; u[b].y += u[b + w].y; }
            float2* dst = a.X + ((size_t)ch * a.nkx + xb0 + q) * a.n_frames + f;
            if (whole) *dst = u[0];
            else { atomicAdd(&dst->x, u[0].x); atomicAdd(&dst->y, u[0].y); }
        }
    }
}

__global__ void __launch_bounds__(TC_CTA_THREADS, 1)
blockdft_tc_kernel(const __grid_constant__ BlockDftTcArgs a) {
    extern __shared__ __align__(16) uint8_t tc_smem_raw[];
    // swizzled operand tiles need a 1024-byte aligned base (the swizzle is a function of address bits)
    // (offset added to the __shared__ array itself, not a round trip through uintptr_t: the compiler then keeps the
    // shared address space and emits STS.128 / LDS.64 instead of generic stores)
    uint8_t* tiles = tc_smem_raw + ((1024u - (tc_smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    float2* S_all = reinterpret_cast<float2*>(tiles + TC_STAGES * TC_STAGE_BYTES);   // epilogue exchange strips, one per group
    uint64_t* bar_a = reinterpret_cast<uint64_t*>(tiles + TC_STAGES * TC_STAGE_BYTES + (TC_EPI_WARPS / 4) * TC_STRIP_BYTES);
    uint64_t* bar_b = bar_a + TC_STAGES;
    uint64_t* bar_m = bar_b + TC_STAGES;
    uint64_t* bar_full = bar_m + TC_STAGES;       // one commit after a tile's last MMA: the accumulators are complete
    uint64_t* bar_empty = bar_full + 1;           // the epilogue warps have read the accumulators out of TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_per_ch = (a.nb + TC_BM - 1) / TC_BM;
    const int n_kc = a.hop / TC_KC;

    // Three mbarrier rings tie the roles together (stage s of chunk number kcg, counted over all tiles of the CTA):
    //   bar_a[s]  one arrival per warp of the producer group: A_hi / A_lo of the stage are written and fenced
    //   bar_b[s]  tx bytes: the B image of the stage has landed
    //   bar_m[s]  tcgen05.commit: every MMA issued so far (in particular those reading stage s) has retired
    if (warp == TC_WARP_B && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            tc_mbar_init(bar_a + s, TC_THREADS / 32 / TC_PG); tc_mbar_init(bar_b + s, 1); tc_mbar_init(bar_m + s, 1);
        }
        tc_mbar_init(bar_full, 1);
        tc_mbar_init(bar_empty, TC_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == TC_WARP_B) {
        // ===== B-operand issuer (one thread): runs up to STAGES chunks ahead of the MMAs' completion, across tiles =====
        if (lane == 0) {
            int kcg = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                const uint8_t* eimg = a.Eimg + (size_t)(tile % a.n_halves) * n_kc * 2 * TC_B_BYTES;
                for (int kb = 0; kb < n_kc; ++kb, ++kcg) {
                    const int sb = kcg % TC_STAGES;
                    // the stage was last read by chunk kcg - STAGES
                    if (kcg >= TC_STAGES) tc_mbar_wait(bar_m + sb, (uint32_t)(kcg / TC_STAGES - 1) & 1);
#ifdef TC_EXPERIMENT_B_DIV      /* timing experiment only (wrong results): fetch 1/DIV of the B bytes */
                    tc_mbar_expect_tx(bar_b + sb, 2 * TC_B_BYTES / TC_EXPERIMENT_B_DIV);
                    tc_bulk_g2s(tiles + sb * TC_STAGE_BYTES + 2 * TC_A_BYTES, eimg + (size_t)kb * 2 * TC_B_BYTES, 2 * TC_B_BYTES / TC_EXPERIMENT_B_DIV, bar_b + sb);
#else
                    tc_mbar_expect_tx(bar_b + sb, 2 * TC_B_BYTES);
                    tc_bulk_g2s(tiles + sb * TC_STAGE_BYTES + 2 * TC_A_BYTES, eimg + (size_t)kb * 2 * TC_B_BYTES, 2 * TC_B_BYTES, bar_b + sb);
#endif
                }
            }
        }
    } else if (warp == TC_WARP_MMA) {
        // ===== MMA issuer (one thread): waits for operands, and at a tile's start for the TMEM to be drained =====
        if (lane == 0) {
            int kcg = 0, it = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
                TC_STAMP(it == 1, 0);
                TC_STAMP(it == 2, 6);
                if (it > 0) {
                    tc_mbar_wait(bar_empty, (uint32_t)(it - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int kc = 0; kc < n_kc; ++kc, ++kcg) {
                    const int s = kcg % TC_STAGES;
                    const uint32_t use = (uint32_t)(kcg / TC_STAGES);
                    tc_mbar_wait(bar_a + s, use & 1);
                    tc_mbar_wait(bar_b + s, use & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    TC_STAMP(it == 1 && kc == 0, 1);
                    const uint32_t sa = tc_smem_u32(tiles + s * TC_STAGE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < TC_EXPERIMENT_KS; ++ks) {
                        const uint64_t d_ahi = tc_desc_sw64(sa + ks * 32);
                        const uint64_t d_alo = tc_desc_sw64(sa + TC_A_BYTES + ks * 32);
                        const uint64_t d_bhi = tc_desc_sw64(sa + 2 * TC_A_BYTES + ks * 32);
                        const uint64_t d_blo = tc_desc_sw64(sa + 2 * TC_A_BYTES + TC_B_BYTES + ks * 32);
                        tc_mma(tmem, d_ahi, d_bhi, (kc | ks) != 0);
                        tc_mma(tmem + TC_BN, d_ahi, d_blo, (kc | ks) != 0);
                        tc_mma(tmem + TC_BN, d_alo, d_bhi, 1);
                    }
                    tc_commit(bar_m + s);
                }
                tc_commit(bar_full);
                TC_STAMP(it == 1, 2);
            }
        }
    } else if (warp < TC_WARP_MMA) {
        // ===== A producers.  TC_PG groups of 256 / TC_PG threads each own every TC_PG-th chunk: a chunk costs its
        // producer a wait + STS + fence.proxy.async + arrive round trip that no amount of look-ahead shortens, so
        // TC_PG chunks are kept in flight side by side instead of one after the other.  The loads of a chunk are
        // issued before the wait for its stage, so they travel while the MMAs that still read it retire =====
        constexpr int PGT = TC_THREADS / TC_PG;                  // threads per producer group
        constexpr int NQ = TC_BM * 4 / PGT;                      // 16-byte pieces per thread and chunk
        constexpr int PF4 = TC_F16 ? 2 : 1;                      // float4 loads per piece (8 half or 4 tf32 samples)
        const int pg = tid / PGT, tg = tid % PGT;
        int a_off[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) a_off[q] = tc_sw64_offset((tg + PGT * q) >> 2, (tg + PGT * q) & 3);
        int it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            const int bid = tile / a.n_halves;
            const int row0 = (bid % tiles_per_ch) * TC_BM;
            const int ch = bid / tiles_per_ch;
            const float* xa = a.x + (long long)ch * a.ch_stride + (long long)a.j0 * a.hop;
            const float4* a_src[NQ];
            float sc[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int p = tg + PGT * q;
                int r = row0 + (p >> 2);
                if (r >= a.nb) r = a.nb - 1;                      // rows past the end: computed, never stored
                a_src[q] = reinterpret_cast<const float4*>(xa + (long long)r * a.hop) + (p & 3) * PF4;
#if TC_F16
                sc[q] = __ldg(a.row_inv + (size_t)ch * a.nb + r);                 // 2^e, inverted below
#else
                sc[q] = 1.f;
#endif
            }
#if TC_F16
#pragma unroll
            for (int q = 0; q < NQ; ++q) sc[q] = __uint_as_float(0x7F000000u - __float_as_uint(sc[q]));         // 1 / 2^e
#endif
            // group pg owns the chunks whose number over all tiles of this CTA is pg (mod TC_PG): always stage pg
            const int kcg0 = it * n_kc;
            for (int kc = ((pg - kcg0) % TC_PG + TC_PG) % TC_PG; kc < n_kc; kc += TC_PG) {
                const int kcg = kcg0 + kc;
                const int s = kcg % TC_STAGES;
                uint8_t* st = tiles + s * TC_STAGE_BYTES;
                float4 cur[NQ][PF4];
#ifndef TC_EXPERIMENT_NO_LDG
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int h = 0; h < PF4; ++h) cur[q][h] = __ldg(a_src[q] + kc * (TC_KC / 4) + h);
#else
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int h = 0; h < PF4; ++h) cur[q][h] = make_float4(0.f, 0.f, 0.f, 0.f);
#endif
                // the MMAs that read this stage (chunk kcg - STAGES) must have retired before it is overwritten
                if (kcg >= TC_STAGES) tc_mbar_wait(bar_m + s, (uint32_t)(kcg / TC_STAGES - 1) & 1);
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
#if TC_F16
                    uint4 hi, lo;
                    tc_split_h2(cur[q][0].x * sc[q], cur[q][0].y * sc[q], hi.x, lo.x);
                    tc_split_h2(cur[q][0].z * sc[q], cur[q][0].w * sc[q], hi.y, lo.y);
                    tc_split_h2(cur[q][1].x * sc[q], cur[q][1].y * sc[q], hi.z, lo.z);
                    tc_split_h2(cur[q][1].z * sc[q], cur[q][1].w * sc[q], hi.w, lo.w);
                    *reinterpret_cast<uint4*>(st + a_off[q]) = hi;
                    *reinterpret_cast<uint4*>(st + TC_A_BYTES + a_off[q]) = lo;
#else
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(cur[q][0].x) & 0xFFFFE000u); lo.x = cur[q][0].x - hi.x;
                    hi.y = __uint_as_float(__float_as_uint(cur[q][0].y) & 0xFFFFE000u); lo.y = cur[q][0].y - hi.y;
                    hi.z = __uint_as_float(__float_as_uint(cur[q][0].z) & 0xFFFFE000u); lo.z = cur[q][0].z - hi.z;
                    hi.w = __uint_as_float(__float_as_uint(cur[q][0].w) & 0xFFFFE000u); lo.w = cur[q][0].w - hi.w;
                    *reinterpret_cast<float4*>(st + a_off[q]) = hi;
                    *reinterpret_cast<float4*>(st + TC_A_BYTES + a_off[q]) = lo;
#endif
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> tensor-core reads
                __syncwarp();                                                      // one arrival per warp
                if (lane == 0) tc_mbar_arrive(bar_a + s);
            }
        }
    } else {
        // ===== epilogue warps: warp w reads TMEM lanes 32 (w % 4) .. +31 = the rows quad * 32 + lane of the tile;
        // main + cross accumulators are added in registers.  The next column group's TMEM loads are issued as soon
        // as this group's values are in the strip; after the last loads have landed the TMEM goes back to the MMAs =====
        const int quad = warp & 3, grp = (warp - TC_WARP_EPI) >> 2;
        const int r = quad * 32 + lane;                                // local row = local index of the frame's last block
        constexpr int NCG = TC_BN / 32 / (TC_EPI_WARPS / 4);           // 32-column groups per epilogue group
        const uint32_t taddr0 = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * NCG * 32);
        float2* S = S_all + grp * (TC_STRIP_BYTES / 8);
        for (int i = (tid - TC_WARP_EPI * 32) & 127; i < TC_SPAD * TC_XROW; i += 128) {             // the zero rows, once
            S[i] = make_float2(0.f, 0.f);
            S[(128 + TC_SPAD) * TC_XROW + i] = make_float2(0.f, 0.f);
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            const int half = tile % a.n_halves;
            const int bid = tile / a.n_halves;
            const int row0 = (bid % tiles_per_ch) * TC_BM;
            const int ch = bid / tiles_per_ch;
            const int row = row0 + r;
            float rs = 1.f;
#if TC_F16
            rs = __ldg(a.row_inv + (size_t)ch * a.nb + (row < a.nb ? row : a.nb - 1)) * a.e_inv;
#endif
            float* qrow = a.X ? nullptr : a.Q + ((size_t)ch * a.nb + (row < a.nb ? row : 0)) * a.qs + half * TC_BN;
            tc_mbar_wait(bar_full, (uint32_t)it & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TC_STAMP(it == 1 && tid == TC_WARP_EPI * 32, 3);
            uint32_t v[32], c[32];
            tc_tmem_ld32(taddr0, v);
            tc_tmem_ld32(taddr0 + TC_BN, c);
#pragma unroll 1
            for (int cg = 0; cg < NCG; ++cg) {
                TC_STAMP(it == 1 && cg == 2 && tid == TC_WARP_EPI * 32, 8);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                TC_STAMP(it == 1 && cg == 2 && tid == TC_WARP_EPI * 32, 9);
                const int gi = half * (TC_BN / 32) + grp * NCG + cg;
                const int B = a.X ? a.gB[gi] : 0;
#ifdef TC_EXPERIMENT_NO_EPI
                const bool live = false;
#else
                const bool live = B > 0;
#endif
                if (a.X) {
                    if (live) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            S[(r + TC_SPAD) * TC_XROW + i] = make_float2(tc_acc(v[2 * i], c[2 * i], rs), tc_acc(v[2 * i + 1], c[2 * i + 1], rs));
                    }
                } else if (row < a.nb) {
                    float4* dst = reinterpret_cast<float4*>(qrow + (grp * NCG + cg) * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        dst[i] = make_float4(tc_acc(v[4 * i], c[4 * i], rs), tc_acc(v[4 * i + 1], c[4 * i + 1], rs),
                                             tc_acc(v[4 * i + 2], c[4 * i + 2], rs), tc_acc(v[4 * i + 3], c[4 * i + 3], rs));
                }
                TC_STAMP(it == 1 && cg == 2 && tid == TC_WARP_EPI * 32, 10);
                if (cg + 1 < NCG) {
                    tc_tmem_ld32(taddr0 + (uint32_t)((cg + 1) * 32), v);
                    tc_tmem_ld32(taddr0 + (uint32_t)(TC_BN + (cg + 1) * 32), c);
                } else {
                    // every accumulator column is in registers or the strip: hand the TMEM back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(bar_empty);
                    TC_STAMP(it == 1 && tid == TC_WARP_EPI * 32, 4);
                }
                if (a.X) {
                    TC_STAMP(it == 1 && cg == 2 && tid == TC_WARP_EPI * 32, 11);
                    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
                    TC_STAMP(it == 1 && cg == 2 && tid == TC_WARP_EPI * 32, 12);
                    if (live) {
                        const int nbin = a.gN[gi], xb0 = a.gX[gi];
                        if (B == 16) tc_frame_sums<16>(a, S, r, row0, ch, nbin, xb0, a.gS[gi]);
                        else if (B == 8) tc_frame_sums<8>(a, S, r, row0, ch, nbin, xb0);
                        else if (B == 4) tc_frame_sums<4>(a, S, r, row0, ch, nbin, xb0);
                        else tc_frame_sums<2>(a, S, r, row0, ch, nbin, xb0);
                    }
                    TC_STAMP(it == 1 && cg == 2 && tid == TC_WARP_EPI * 32, 13);
                    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
                    TC_STAMP(it == 1 && cg == 2 && tid == TC_WARP_EPI * 32, 14);
                }
            }
            TC_STAMP(it == 1 && tid == TC_WARP_EPI * 32, 5);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
}

inline size_t blockdft_tc_smem_bytes() { return (size_t)TC_STAGES * TC_STAGE_BYTES + (TC_EPI_WARPS / 4) * TC_STRIP_BYTES + (3 * TC_STAGES + 2) * 8 + 16 + 1024; }

}  // namespace o4
