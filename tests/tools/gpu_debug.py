"""Developer script: component-by-component error report of the CUDA path vs the oracle/goldens,
plus a rough timing.  Run on the GPU box: `python tools/gpu_debug.py`."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "audio-analyzer-omega_b200"))

from omega4_b200.plan import AnalysisPlan, BASELINE_CONFIGS, rfft_batch_host  # noqa: E402
from omega4_b200 import _native as N  # noqa: E402
from oracle import oracle_np as O  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def dberr(got, ref, floor_db=-60):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    peak = ref.max(axis=-1, keepdims=True)
    sig = ref >= np.maximum(peak * 10 ** (floor_db / 20), 1e-30)
    if not sig.any():
        return 0.0
    return float(np.abs(20 * np.log10(np.maximum(got[sig], 1e-30)) - 20 * np.log10(ref[sig])).max())


def main():
    rng = np.random.default_rng(0)
    print("== rfft_batch vs float64 numpy")
    for n in (512, 1024, 2048, 4096, 8192, 16384, 32768):
        x = rng.standard_normal((5, n)).astype(np.float32)
        w = np.hanning(n).astype(np.float32)
        try:
            mag, cx = rfft_batch_host(x, w)
            ref = np.fft.rfft(x.astype(np.float64) * w, axis=1)
            print(n, "max|dX|/max|X| =", np.abs(cx - ref).max() / np.abs(ref).max(),
                  " mag rel =", np.abs(mag - np.abs(ref)).max() / np.abs(ref).max())
        except Exception as e:
            print(n, "FAILED", e)

    print("== multires baseline golden")
    g = np.load(os.path.join(G, "multires_baseline.npz"))
    plan = AnalysisPlan(48000, BASELINE_CONFIGS, 512)
    out = plan.analyze_host(g["x"][None, :], want_combined=True, want_magnitudes=True, want_meters=True, want_series=True)
    comb = out["combined"][0]
    print("combined dB err (>-60dB):", dberr(comb, g["combined"]), " zero pattern equal:",
          np.array_equal(comb == 0, g["combined"] == 0))
    for k in (15, 16, 50, 95):
        for r in range(4):
            key = f"mag_h{k}_r{r}"
            if key in g:
                print(key, "dB err", dberr(out["magnitudes"][r][0, k], g[key]),
                      "abs err/peak", np.abs(out["magnitudes"][r][0, k] - g[key]).max() / g[key].max())
    print("== meters stream golden")
    gm = np.load(os.path.join(G, "meters_stream.npz"))
    o2 = plan.analyze_host(gm["x"][None, :], want_combined=False, want_meters=True, want_series=True)
    f = int(gm["first_hop"])
    print("lufs_inst max err", np.abs(o2["lufs_inst"][0, f:] - gm["lufs_inst"]).max())
    print("tp_db max err", np.abs(o2["tp_db"][0, f:] - gm["tp_db"]).max())
    print("meters max err per column", np.abs(o2["meters"][0, f:] - gm["meters"]).max(axis=0))
    print("meters before first:", o2["meters"][0, :f])
    print("== meters stats golden (frames mode)")
    gs = np.load(os.path.join(G, "meters_stats.npz"))
    frames = gs["gains"][:, None] * gs["base"][np.arange(len(gs["gains"])) % 4]
    li, tp, _ = plan.meter_frames_host(frames)
    print("frames lufs err", np.abs(li - gs["lufs_inst"]).max(), " tp err", np.abs(tp - gs["tp_db"]).max())
    st = np.zeros((1, N.METER_STATE_DOUBLES))
    m = plan.meter_stats_host(gs["lufs_inst"], gs["tp_db"], state=st, fresh=True)
    print("stats err per column (exact series in)", np.abs(m[0] - gs["meters"]).max(axis=0))
    # tiled with carry
    st = np.zeros((1, N.METER_STATE_DOUBLES))
    parts = []
    for s in range(0, 4200, 700):
        parts.append(plan.meter_stats_host(gs["lufs_inst"][s:s + 700], gs["tp_db"][s:s + 700], state=st, fresh=(s == 0))[0])
    print("stats err tiled", np.abs(np.concatenate(parts) - gs["meters"]).max(axis=0))

    print("== timing (device resident)")
    import torch
    from omega4_b200.batch.driver import device_synth
    n_ch, secs = 256, 20
    n_hops = secs * 48000 // 512
    x = device_synth(n_ch // 2, 2, n_hops * 512, 48000)
    comb = torch.empty((n_ch, n_hops, 512), dtype=torch.float32, device="cuda")
    met = torch.empty((n_ch, n_hops, 5), dtype=torch.float32, device="cuda")
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        plan.analyze_device(x, n_hops, 0, combined=comb, meters=met, flags=N.FLAG_TIME_KERNELS | N.FLAG_FRESH_METERS | N.FLAG_CONCURRENT_METERS)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"iter {it}: {dt*1e3:.1f} ms  -> {n_ch*n_hops/dt/1e6:.2f} M channel-hops/s, "
              f"{n_ch/2*n_hops*512/48000/dt:.0f} stereo stream-s/s")
    for name, ms in plan.kernel_times():
        print(f"   {name:20s} {ms:8.3f} ms (concurrent)")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    plan.analyze_device(x, n_hops, 0, combined=comb, meters=met, flags=N.FLAG_TIME_KERNELS | N.FLAG_FRESH_METERS)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"serial: {dt*1e3:.1f} ms")
    for name, ms in plan.kernel_times():
        print(f"   {name:20s} {ms:8.3f} ms (serial)")
    print("meters sample", met[0, -1].cpu().numpy(), "comb finite", bool(torch.isfinite(comb).all()))


if __name__ == "__main__":
    main()
