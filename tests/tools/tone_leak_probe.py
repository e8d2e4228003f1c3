"""Strong low-frequency tone next to the few-bin resolutions: absolute error of each spectrum path in units of the
largest magnitude of the transform the value came from (float64 evaluation of the same frames as truth).
    python tests/tools/tone_leak_probe.py            (GPU box)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "audio-analyzer-omega_b200"))
from omega4_b200 import _native as N
from omega4_b200.plan import AnalysisPlan, BASELINE_CONFIGS, CONFIG5_96K
from oracle import oracle_np as O

HOP = 512


def truth_rows(x, sr, configs, hops):
    """float64 evaluation of the combined rows + per-resolution transform maxima at the given hops."""
    mr = O.OracleMultiResFFT(sr, 20000, list(configs))
    rows, tmax, rows32 = {}, {}, {}
    for k in hops:
        e = (k + 1) * HOP
        res64, res32, mx = {}, {}, {}
        for i, c in enumerate(mr.configs):
            if e < c.fft_size:
                continue
            fr = x[e - c.fft_size:e]
            m64 = np.abs(np.fft.rfft(fr.astype(np.float64) * mr.windows[i].astype(np.float64))) * mr.bin_weights(i)
            res64[i] = m64
            res32[i] = (np.abs(np.fft.rfft(fr * mr.windows[i])) * mr.bin_weights(i)).astype(np.float32)
            mx[i] = m64.max()
        rows[k] = mr.combine({i: v.astype(np.float32) for i, v in res64.items()}, 512)[0].astype(np.float64)
        rows32[k] = mr.combine(res32, 512)[0].astype(np.float64)
        tmax[k] = mx
    return rows, rows32, tmax


def probe(name, sr, plan_cfg, oracle_cfg, x, hops, bins_of_res):
    rows, rows32, tmax = truth_rows(x, sr, oracle_cfg, hops)
    plan = AnalysisPlan(sr, plan_cfg, 512)
    print(f"== {name}")
    for tag, fl in (("tensor-core hop-block DFT", 0), ("full FFT kernel", N.FLAG_NO_BLOCKDFT)):
        comb = plan.analyze_host(x[None, :], want_meters=False, flags=fl)["combined"][0].astype(np.float64)
        for r, bins in bins_of_res.items():
            worst, worst_ref, wdb, wat = 0.0, 0.0, 0.0, None
            for k in hops:
                if r not in tmax[k]:
                    continue
                err = np.abs(comb[k, bins] - rows[k][bins]) / tmax[k][r]
                err_ref = np.abs(rows32[k][bins] - rows[k][bins]) / tmax[k][r]
                j = int(np.argmax(err))
                if err[j] > worst:
                    worst, wat = err[j], (k, bins[j])
                    wdb = 20 * np.log10(rows[k][bins[j]] / tmax[k][r])
                worst_ref = max(worst_ref, err_ref.max())
            print(f"  {tag:28s} res {r}: worst |err| / transform max {worst:.2e} (value at {wdb:.0f} dB, hop/bin {wat}); "
                  f"numpy float32 path: {worst_ref:.2e}")
    plan.close()


g = np.load(os.path.join(ROOT, "tests", "golden", "multires_96k_stress.npz"))
probe("96 kHz, 32768 / 16384 (click, silence, 0.95 x 41 Hz + 0.3 x 130 Hz)", 96000, CONFIG5_96K, O.CONFIG5_96K, g["x"],
      list(range(64, 150, 5)) + [149], {0: [1], 1: [2, 3, 4, 5]})
n = 60 * HOP
t = np.arange(n) / 48000.0
rng = np.random.default_rng(1)
x = (0.95 * np.sin(2 * np.pi * 41.0 * t) + 0.3 * np.sin(2 * np.pi * 130.0 * t) + 1e-4 * rng.standard_normal(n)).astype(np.float32)
probe("48 kHz, 8192 / 4096 (0.95 x 41 Hz + 0.3 x 130 Hz)", 48000, BASELINE_CONFIGS, O.BASELINE_CONFIGS, x,
      list(range(16, 60, 4)), {0: [1, 2, 3, 4, 5], 1: list(range(6, 26))})
x = (0.95 * np.sin(2 * np.pi * 300.0 * t) + 1e-4 * rng.standard_normal(n)).astype(np.float32)
probe("48 kHz, 8192 / 4096 (0.95 x 300 Hz)", 48000, BASELINE_CONFIGS, O.BASELINE_CONFIGS, x,
      list(range(16, 60, 4)), {0: [1, 2, 3, 4, 5], 1: list(range(6, 26))})
