"""Worst-case accuracy probe of the three spectrum paths: strong pure tones inside / next to the sparse
low-frequency bins, float64 reference.  Prints the max dB error over bins within 60 dB of the row max and
the max absolute error of the quieter bins relative to the 0.01 dB-at-the-floor allowance."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "audio-analyzer-omega_b200"))
from omega4_b200 import _native as N, tables
from omega4_b200.plan import AnalysisPlan, BASELINE_CONFIGS

HOP = 512
p = AnalysisPlan(48000, BASELINE_CONFIGS, 512)
n = 40 * HOP
t = np.arange(n) / 48000.0
cases = {"58.6 Hz (8192 bin 10, between needed bins)": [(58.59375, 0.9)],
         "41 Hz + 300 Hz": [(41.0, 0.9), (300.0, 0.5)],
         "1 kHz strong + 97 Hz -50 dB": [(1000.0, 0.9), (97.0, 0.9 * 10 ** (-50 / 20))],
         "250.5 Hz (4096 region)": [(250.5, 0.9)],
         "sum of 30 tones 20-1000 Hz": [(f, 0.03) for f in np.linspace(23.0, 990.0, 30)]}
for name, tones in cases.items():
    x = sum(a * np.sin(2 * np.pi * f * t + 0.1 * i) for i, (f, a) in enumerate(tones)).astype(np.float32)
    mags = []
    for (fr, nfft, _h, w, wt) in BASELINE_CONFIGS:
        frame = x[n - nfft:].astype(np.float64) * np.blackman(nfft).astype(np.float32).astype(np.float64)
        freqs = np.fft.rfftfreq(nfft, 1 / 48000)
        mags.append((np.abs(np.fft.rfft(frame)) * tables.psycho_weights(freqs, fr, w)).astype(np.float64))
    ct = tables.combine_tables(48000, 20000, [c[1] for c in BASELINE_CONFIGS], [c[0] for c in BASELINE_CONFIGS], 512)
    want = np.zeros(512)
    for m, (idx, lo, frac) in zip(mags, ct):
        want[idx] = m[lo] + (m[lo + 1] - m[lo]) * frac.astype(np.float64)
    peak = want.max(); floor = peak * 1e-3
    sig = want >= floor
    line = f"{name:44s}"
    for label, fl in (("tc", 0), ("fp32", N.FLAG_NO_TENSOR), ("fft", N.FLAG_NO_BLOCKDFT)):
        got = p.analyze_host(x[None, :], want_meters=False, flags=fl)["combined"][0, 39].astype(np.float64)
        e = np.abs(20 * np.log10(np.maximum(got[sig], 1e-30)) - 20 * np.log10(want[sig])).max()
        q = (np.abs(got - want)[~sig] / (floor * (10 ** (0.01 / 20) - 1))).max() if (~sig).any() else 0.0
        line += f" | {label}: {e:.1e} dB, quiet {q:.2f}x allowance"
    print(line)
