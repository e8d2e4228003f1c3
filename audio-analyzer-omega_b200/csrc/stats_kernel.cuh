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
constexpr int ST_TILE = 128;    // hops of the input series staged in shared memory at a time
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

// first index i in [0, n) with s[i] >= v (n if none), found by one warp with a 32-ary search
__device__ __forceinline__ int warp_lower_bound(const double* s, int n, double v, int lane) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int step = (hi - lo + 31) >> 5;
        const int idx = lo + lane * step;
        const bool less = (idx < hi) && (s[idx] < v);
        const int c = __popc(__ballot_sync(0xffffffffu, less));
        if (step == 1) return lo + c;
        if (c == 0) return lo;
        const int nhi = min(hi, lo + c * step);
        lo = lo + (c - 1) * step + 1;
        hi = nhi;
    }
    return lo;
}

// Per hop three block barriers:
//   [shift-read] | [shift-write + ring / running-sum update by thread 0] |
//   [warps 0,1: search positions for the NEXT hop] || [warp 2: peak max, warp 3: percentiles + output]
__global__ void __launch_bounds__(ST_THREADS)
stats_kernel(const __grid_constant__ StatsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* win = reinterpret_cast<double*>(smem_raw);        // ring of the last 3600 values
    double* srt = win + ST_I;                                 // sorted gated values (ST_SORT slots)
    double* pkw = srt + ST_SORT;                              // ring of the last 60 peaks
    __shared__ int sh_i[8];       // 2 ns (current), 3 n_hist, 4 head, 5 n_tp, 7 pk_head
    __shared__ double sh_d[8];    // 1 sum_i, 2 M, 3 S
    __shared__ int sh_prem[2], sh_pins[2];    // positions for hop k live in slot k & 1
    __shared__ double sh_nv[2];
    // the per-frame series are staged ST_TILE hops at a time: no global-load latency inside the hop loop
    __shared__ double sh_lufs[ST_TILE + 1], sh_tp[ST_TILE];

    const int ch = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    double* st = a.state + (size_t)ch * ST_STATE;
    const double* lufs = a.lufs + (size_t)ch * a.n_frames;
    const double* tps = a.tp + (size_t)ch * a.n_frames;

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
    // thread-0 bookkeeping
    int head = 0, pk_head = 0, ns = 0, cnt_m = 0, cnt_s = 0, ns_inv_for = -1;
    double sum_i = 0.0, sum_m = 0.0, sum_s = 0.0, inv_m = 1.0, inv_s = 1.0, inv_ns = 0.0;
    if (tid == 0) {
        for (int i = 0; i < n_hist; ++i) if (win[i] > a.gate) { ++ns; sum_i += win[i]; }
        cnt_m = n_hist < ST_M ? n_hist : ST_M;
        cnt_s = n_hist < ST_S ? n_hist : ST_S;
        for (int i = n_hist - cnt_m; i < n_hist; ++i) sum_m += win[i];
        for (int i = n_hist - cnt_s; i < n_hist; ++i) sum_s += win[i];
        if (cnt_m > 0) inv_m = 1.0 / (double)cnt_m;
        if (cnt_s > 0) inv_s = 1.0 / (double)cnt_s;
        sh_i[2] = ns; sh_i[3] = n_hist; sh_i[4] = 0; sh_i[5] = n_tp;
    }
    // hops before the first meter frame keep the current outputs
    const int kbeg = a.first_frame < a.n_frames ? a.first_frame : a.n_frames;
    for (int k = tid; k < kbeg; k += ST_THREADS) {
        float* orow = a.out + ((size_t)ch * a.n_frames + k) * 5;
#pragma unroll
        for (int i = 0; i < 5; ++i) orow[i] = (float)cur[i];
    }
    __syncthreads();
    // stage the first tile of the series
    for (int i = tid; i <= ST_TILE; i += ST_THREADS) {
        const int k = kbeg + i;
        if (k < a.n_frames) { sh_lufs[i] = lufs[k]; if (i < ST_TILE) sh_tp[i] = tps[k]; }
    }
    __syncthreads();
    // positions for the first hop
    if (kbeg < a.n_frames) {
        const double nv = sh_lufs[0];
        const int nh = sh_i[3], nsv = sh_i[2];
        if (warp == 0) {
            int prem = -1;
            if (nh == ST_I) { double ov = win[sh_i[4]]; if (ov > a.gate) prem = warp_lower_bound(srt, nsv, ov, lane); }
            if (lane == 0) sh_prem[kbeg & 1] = prem;
        } else if (warp == 1) {
            int pins = (nv > a.gate) ? warp_lower_bound(srt, nsv, nv, lane) : -1;
            if (lane == 0) { sh_pins[kbeg & 1] = pins; sh_nv[kbeg & 1] = nv; }
        }
    }
    __syncthreads();

    for (int k = kbeg; k < a.n_frames; ++k) {
        const int kt = (k - kbeg) % ST_TILE;          // position inside the staged tile
        if (kt == 0 && k != kbeg) {
            // restage: sh_lufs[0..ST_TILE] = lufs[k .. k+ST_TILE], sh_tp likewise (all readers of the
            // previous tile are past the barrier that ended the previous iteration)
            for (int i = tid; i <= ST_TILE; i += ST_THREADS) {
                const int kk = k + i;
                if (kk < a.n_frames) { sh_lufs[i] = lufs[kk]; if (i < ST_TILE) sh_tp[i] = tps[kk]; }
            }
            __syncthreads();
        }
        // ---- shift of the sorted array: remove old index prem (if >= 0), insert before old index pins (if >= 0)
        const int prem = sh_prem[k & 1], pins = sh_pins[k & 1], ns_old = sh_i[2];
        int lo = ns_old, hi = -1;
        if (prem >= 0 && pins >= 0) {
            if (pins > prem) { lo = prem + 1; hi = pins - 1; }        // shift left by one
            else { lo = pins; hi = prem - 1; }                        // shift right by one
        } else if (prem >= 0) { lo = prem + 1; hi = ns_old - 1; }
        else if (pins >= 0) { lo = pins; hi = ns_old - 1; }
        const int delta = (prem >= 0 && (pins < 0 || pins > prem)) ? -1 : +1;
        const int cnt = hi - lo + 1;
        constexpr int PER = (ST_I + ST_THREADS - 1) / ST_THREADS;     // fully unrolled -> registers
        double tmp[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            int idx = lo + tid + i * ST_THREADS;
            if (i * ST_THREADS < cnt && idx <= hi) tmp[i] = srt[idx];
        }
        const double nv = sh_nv[k & 1];
        const double np_ = sh_tp[kt];
        const double nv_next = (k + 1 < a.n_frames) ? sh_lufs[kt + 1] : 0.0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            int idx = lo + tid + i * ST_THREADS;
            if (i * ST_THREADS < cnt && idx <= hi) srt[idx + delta] = tmp[i];
        }
        if (tid == 0) {
            if (pins >= 0) srt[pins - ((prem >= 0 && pins > prem) ? 1 : 0)] = nv;
            // ring of instantaneous values
            int slot;
            double ov = 0.0;
            const bool has_old = (n_hist == ST_I);
            if (has_old) { slot = head; ov = win[head]; head = (head + 1 == ST_I) ? 0 : head + 1; }
            else { slot = n_hist; ++n_hist; }
            win[slot] = nv;
            if (prem >= 0) sum_i -= ov;
            if (pins >= 0) sum_i += nv;
            ns = ns_old + (pins >= 0 ? 1 : 0) - (prem >= 0 ? 1 : 0);
            if (ns == 0) sum_i = 0.0;
            // running sums of the last 24 / 180 values (newest at `slot`)
            sum_m += nv;
            if (cnt_m == ST_M) { int j = slot - ST_M; if (j < 0) j += ST_I; sum_m -= win[j]; }
            else { ++cnt_m; inv_m = 1.0 / (double)cnt_m; }
            sum_s += nv;
            if (cnt_s == ST_S) { int j = slot - ST_S; if (j < 0) j += ST_I; sum_s -= win[j]; }
            else { ++cnt_s; inv_s = 1.0 / (double)cnt_s; }
            if (ns != ns_inv_for) { inv_ns = ns > 0 ? 1.0 / (double)ns : 0.0; ns_inv_for = ns; }
            // ring of peaks
            int pslot;
            if (n_tp == ST_P) { pslot = pk_head; pk_head = (pk_head + 1 == ST_P) ? 0 : pk_head + 1; }
            else { pslot = n_tp; ++n_tp; }
            pkw[pslot] = np_;
            sh_i[2] = ns; sh_i[3] = n_hist; sh_i[4] = head; sh_i[5] = n_tp;
            sh_d[1] = sum_i * inv_ns; sh_d[2] = sum_m * inv_m; sh_d[3] = sum_s * inv_s;
        }
        __syncthreads();
        // ---- read-only phase: searches for hop k+1 || statistics of hop k
        {
            const int nh = sh_i[3], nsv = sh_i[2];
            if (warp == 0) {
                int prem_n = -1;
                if (k + 1 < a.n_frames && nh == ST_I) {
                    double ov = win[sh_i[4]];
                    if (ov > a.gate) prem_n = warp_lower_bound(srt, nsv, ov, lane);
                }
                if (lane == 0) sh_prem[(k + 1) & 1] = prem_n;
            } else if (warp == 1) {
                int pins_n = (k + 1 < a.n_frames && nv_next > a.gate) ? warp_lower_bound(srt, nsv, nv_next, lane) : -1;
                if (lane == 0) { sh_pins[(k + 1) & 1] = pins_n; sh_nv[(k + 1) & 1] = nv_next; }
            } else if (warp == 2) {
                const int ntp = sh_i[5];
                double v = -CUDART_INF;
                for (int i = lane; i < ntp; i += 32) v = fmax(v, pkw[i]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
                // lane 0: P95, lane 1: P10
                double pv = (nsv > 0 && lane < 2) ? st_percentile(srt, nsv, lane == 0 ? 0.95 : 0.10) : 0.0;
                double p10 = __shfl_sync(0xffffffffu, pv, 1);
                if (lane == 0) {
                    float* orow = a.out + ((size_t)ch * a.n_frames + k) * 5;
                    cur[0] = sh_d[2]; cur[1] = sh_d[3];
                    cur[2] = (nsv > 0) ? sh_d[1] : -100.0;
                    cur[3] = (nsv > 0) ? pv - p10 : 0.0;
                    cur[4] = v;
                    orow[0] = (float)cur[0]; orow[1] = (float)cur[1]; orow[2] = (float)cur[2];
                    orow[3] = (float)cur[3]; orow[4] = (float)cur[4];
                }
            }
        }
        // the barrier after the next iteration's shift-read orders this phase's reads of srt/win/pkw
        // before the next writes; the positions for hop k+1 are published by the barrier below
        __syncthreads();
    }

    // ---- write state back (chronological order)
    if (tid == 0) { sh_i[3] = n_hist; sh_i[4] = head; sh_i[5] = n_tp; sh_i[7] = pk_head; }
    __syncthreads();
    {
        const int nh = sh_i[3], hd = sh_i[4], ntp = sh_i[5], ph = sh_i[7];
        for (int i = tid; i < nh; i += ST_THREADS) st[8 + i] = win[(nh == ST_I) ? (hd + i) % ST_I : i];
        for (int i = tid; i < ntp; i += ST_THREADS) st[8 + ST_I + i] = pkw[(ntp == ST_P) ? (ph + i) % ST_P : i];
        if (tid == 0) { st[0] = (double)nh; st[1] = (double)ntp; st[7] = 1.0; }
        if (tid == 64) { st[2] = cur[0]; st[3] = cur[1]; st[4] = cur[2]; st[5] = cur[3]; st[6] = cur[4]; }
    }
}

inline size_t stats_smem_bytes() { return (size_t)(ST_I + ST_SORT + ST_P + 4) * sizeof(double); }

}  // namespace o4
