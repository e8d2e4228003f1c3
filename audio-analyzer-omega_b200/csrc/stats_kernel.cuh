// stats_kernel.cuh -- the deque statistics of ProfessionalMetering.calculate_lufs.
//
// Replaces omega4/panels/professional_meters.py:248-279 for a series of per-frame values:
//   momentary  = mean of the last 24 instantaneous-LUFS values (dB domain, -100 entries included)
//   short_term = mean of the last 180
//   integrated = mean of the values > -70 among the last 3600, else -100
//   range      = P95 - P10 (numpy 'linear' percentile) of those gated values, else 0
//   true_peak  = max of the last 60 per-frame true peaks
// One CTA per channel walks the hops in order (the only sequential dependency of the whole
// path).  The window of the last 3600 values and a SORTED copy of its gated members live in
// shared memory; every hop removes the expiring value and inserts the new one by a parallel
// shift, so the exact order statistics numpy would compute are read directly.  State (the
// window, the peak window, the current outputs) is carried in global memory between calls so
// long streams can be processed in time tiles and the streaming shim can push one hop at a time.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace o4 {

constexpr int ST_M = 24;       // int(0.4 * 60)
constexpr int ST_S = 180;      // int(3.0 * 60)
constexpr int ST_I = 3600;     // int(60 * 60)
constexpr int ST_P = 60;       // int(1.0 * 60)
constexpr int ST_SORT = 4096;  // padded sort size
constexpr int ST_THREADS = 128;
// per-channel state, in doubles: [0] n_hist, [1] n_tp, [2..6] current outputs, [8 .. 8+3600) window
// (oldest first), [8+3600 .. 8+3600+60) peak window (oldest first)
constexpr int ST_STATE = 8 + ST_I + ST_P;

struct StatsArgs {
    const double* lufs;        // [n_ch][n_frames]
    const double* tp;          // [n_ch][n_frames]
    int n_ch;
    int n_frames;
    int first_frame;           // hops before this do not push (no meter frame yet)
    double gate;               // -70
    double* state;             // [n_ch][ST_STATE], zero-initialised = fresh meters
    float* out;                // [n_ch][n_frames][5]  M, S, I, LRA, TP
    int fresh;                 // 1: ignore state contents on entry (treat as new meters)
};

__device__ __forceinline__ double st_percentile(const double* s, int n, double q) {
    // numpy.percentile(..., method='linear') on an ascending array of n >= 1 values
    double vi = (double)n * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0;
    if (vi < 0.0) vi = 0.0;
    if (vi > (double)(n - 1)) vi = (double)(n - 1);
    double fl = floor(vi);
    int lo = (int)fl;
    int hi = lo + 1 < n ? lo + 1 : n - 1;
    double t = vi - fl;
    double av = s[lo], bv = s[hi];
    double d = bv - av;
    return (t >= 0.5) ? (bv - d * (1.0 - t)) : (av + d * t);
}

__global__ void __launch_bounds__(ST_THREADS)
stats_kernel(const __grid_constant__ StatsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* win = reinterpret_cast<double*>(smem_raw);        // ring of the last 3600 values
    double* srt = win + ST_I;                                 // sorted gated values (ST_SORT slots)
    double* pkw = srt + ST_SORT;                              // ring of the last 60 peaks
    __shared__ int sh_i[8];
    __shared__ double sh_d[8];

    const int ch = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    double* st = a.state + (size_t)ch * ST_STATE;

    int n_hist = 0, n_tp = 0;
    double cur[5] = {-100.0, -100.0, -100.0, 0.0, -100.0};
    if (!a.fresh) {
        n_hist = (int)st[0];
        n_tp = (int)st[1];
        if (n_hist > 0 || n_tp > 0 || st[7] != 0.0) {
#pragma unroll
            for (int i = 0; i < 5; ++i) cur[i] = st[2 + i];
        }
    }
    for (int i = tid; i < n_hist; i += ST_THREADS) win[i] = st[8 + i];
    for (int i = tid; i < n_tp; i += ST_THREADS) pkw[i] = st[8 + ST_I + i];
    __syncthreads();
    // sorted gated copy: stable compaction is not needed, only the multiset
    for (int i = tid; i < ST_SORT; i += ST_THREADS) {
        double v = (i < n_hist) ? win[i] : CUDART_INF;
        srt[i] = (v > a.gate) ? v : CUDART_INF;               // non-gated entries sort to the end
    }
    __syncthreads();
    for (int k = 2; k <= ST_SORT; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < ST_SORT; i += ST_THREADS) {
                int ixj = i ^ j;
                if (ixj > i) {
                    double x = srt[i], y = srt[ixj];
                    bool up = ((i & k) == 0);
                    if ((x > y) == up) { srt[i] = y; srt[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
    // thread-0 bookkeeping (replicated in registers of thread 0 only)
    int head = 0;                 // index of the oldest value in win (ring)
    int pk_head = 0;
    int ns = 0;                   // number of gated values in srt
    double sum_i = 0.0;           // sum of gated values
    if (tid == 0) {
        for (int i = 0; i < n_hist; ++i) if (win[i] > a.gate) { ++ns; sum_i += win[i]; }
        // ring invariants: when n_hist == ST_I the oldest is at head and new values overwrite it
    }
    __syncthreads();

    for (int k = 0; k < a.n_frames; ++k) {
        float* orow = a.out + ((size_t)ch * a.n_frames + k) * 5;
        if (k < a.first_frame) {                              // block-uniform
            if (tid == 0) { orow[0] = (float)cur[0]; orow[1] = (float)cur[1]; orow[2] = (float)cur[2];
                            orow[3] = (float)cur[3]; orow[4] = (float)cur[4]; }
            continue;
        }
        const double nv = a.lufs[(size_t)ch * a.n_frames + k];
        const double np_ = a.tp[(size_t)ch * a.n_frames + k];
        // ---- phase A: thread 0 updates rings and finds remove / insert positions in srt
        if (tid == 0) {
            double ov = 0.0;
            bool has_old = (n_hist == ST_I);
            int slot;
            if (has_old) { slot = head; ov = win[head]; head = (head + 1) % ST_I; }
            else { slot = (head + n_hist) % ST_I; ++n_hist; }
            win[slot] = nv;
            int pslot;
            if (n_tp == ST_P) { pslot = pk_head; pk_head = (pk_head + 1) % ST_P; }
            else { pslot = (pk_head + n_tp) % ST_P; ++n_tp; }
            pkw[pslot] = np_;
            // positions in the sorted gated array
            int prem = -1, pins = -1;
            bool rem = has_old && (ov > a.gate);
            bool ins = (nv > a.gate);
            if (rem) {                                        // first index with srt[idx] >= ov
                int lo = 0, hi = ns;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (srt[mid] < ov) lo = mid + 1; else hi = mid; }
                prem = lo;
                sum_i -= ov;
            }
            if (ins) {                                        // insertion point in the array AFTER removal
                int lo = 0, hi = ns;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (srt[mid] < nv) lo = mid + 1; else hi = mid; }
                pins = lo;                                    // position in the current (pre-removal) array
                sum_i += nv;
            }
            sh_i[0] = prem; sh_i[1] = pins; sh_i[2] = ns; sh_i[3] = n_hist; sh_i[4] = head;
            sh_i[5] = n_tp;
            sh_d[0] = nv;
            ns += (ins ? 1 : 0) - (rem ? 1 : 0);
            sh_i[6] = ns;
            if (ns == 0) sum_i = 0.0;                         // kill drift when the gate empties the set
            sh_d[1] = sum_i;
        }
        __syncthreads();
        // ---- phase B: parallel shift of srt
        {
            const int prem = sh_i[0], pins = sh_i[1], ns_old = sh_i[2];
            // new array = old with element prem removed (if >= 0) and nv inserted before old index pins (if >= 0)
            // Element at old index i moves to i - (prem >= 0 && i > prem) + (pins >= 0 && i >= pins).
            // Only indices between min and max of the two positions change place.
            int lo = ns_old, hi = -1;                         // affected old-index range [lo, hi]
            if (prem >= 0 && pins >= 0) {
                if (pins > prem) { lo = prem + 1; hi = pins - 1; }        // shift left by one
                else { lo = pins; hi = prem - 1; }                        // shift right by one
            } else if (prem >= 0) { lo = prem + 1; hi = ns_old - 1; }     // shift left
            else if (pins >= 0) { lo = pins; hi = ns_old - 1; }           // shift right
            const int delta = (prem >= 0 && (pins < 0 || pins > prem)) ? -1 : +1;
            const int cnt = hi - lo + 1;
            constexpr int PER = (ST_I + ST_THREADS - 1) / ST_THREADS;     // 29, fully unrolled -> registers
            double tmp[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                int idx = lo + tid + i * ST_THREADS;
                if (i * ST_THREADS < cnt && idx <= hi) tmp[i] = srt[idx];
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                int idx = lo + tid + i * ST_THREADS;
                if (i * ST_THREADS < cnt && idx <= hi) srt[idx + delta] = tmp[i];
            }
            if (tid == 0 && pins >= 0) {
                int dst = pins - ((prem >= 0 && pins > prem) ? 1 : 0);
                srt[dst] = sh_d[0];
            }
        }
        __syncthreads();
        // ---- phase C: read-only statistics, one warp each
        {
            const int nh = sh_i[3], hd = sh_i[4], ntp = sh_i[5], nsn = sh_i[6];
            // chronological index c (0 = oldest) lives at ring slot (hd + c) % ST_I when nh == ST_I,
            // else at slot c (head stays 0 until the ring is full).
            if (warp == 0) {
                int cnt = nh < ST_M ? nh : ST_M;
                double v = 0.0;
                if (lane < cnt) { int c = nh - cnt + lane; v = win[(nh == ST_I) ? (hd + c) % ST_I : c]; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) sh_d[2] = v / (double)cnt;
            } else if (warp == 1) {
                int cnt = nh < ST_S ? nh : ST_S;
                double v = 0.0;
                for (int i = lane; i < cnt; i += 32) { int c = nh - cnt + i; v += win[(nh == ST_I) ? (hd + c) % ST_I : c]; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) sh_d[3] = v / (double)cnt;
            } else if (warp == 2) {
                double v = -CUDART_INF;
                for (int i = lane; i < ntp; i += 32) v = fmax(v, pkw[i]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
                if (lane == 0) sh_d[4] = v;
            } else if (warp == 3 && lane == 0) {
                if (nsn > 0) {
                    sh_d[5] = sh_d[1] / (double)nsn;
                    sh_d[6] = st_percentile(srt, nsn, 0.95) - st_percentile(srt, nsn, 0.10);
                } else { sh_d[5] = -100.0; sh_d[6] = 0.0; }
            }
        }
        __syncthreads();
        cur[0] = sh_d[2]; cur[1] = sh_d[3]; cur[2] = sh_d[5]; cur[3] = sh_d[6]; cur[4] = sh_d[4];
        if (tid == 0) { orow[0] = (float)cur[0]; orow[1] = (float)cur[1]; orow[2] = (float)cur[2];
                        orow[3] = (float)cur[3]; orow[4] = (float)cur[4]; }
        // the barrier at the top of the next iteration's phase A->B orders these reads before the next writes
        __syncthreads();
    }

    // ---- write state back (chronological order)
    if (tid == 0) { sh_i[3] = n_hist; sh_i[4] = head; sh_i[5] = n_tp; sh_i[7] = pk_head; }
    __syncthreads();
    {
        const int nh = sh_i[3], hd = sh_i[4], ntp = sh_i[5], ph = sh_i[7];
        // reading and writing different memories (smem -> global): no hazard
        for (int i = tid; i < nh; i += ST_THREADS) st[8 + i] = win[(nh == ST_I) ? (hd + i) % ST_I : i];
        for (int i = tid; i < ntp; i += ST_THREADS) st[8 + ST_I + i] = pkw[(ntp == ST_P) ? (ph + i) % ST_P : i];
        if (tid == 0) { st[0] = (double)nh; st[1] = (double)ntp; st[7] = 1.0; }
        if (tid == 0) { st[2] = cur[0]; st[3] = cur[1]; st[4] = cur[2]; st[5] = cur[3]; st[6] = cur[4]; }
    }
}

inline size_t stats_smem_bytes() { return (size_t)(ST_I + ST_SORT + ST_P + 4) * sizeof(double); }

}  // namespace o4
