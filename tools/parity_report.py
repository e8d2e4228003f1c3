"""Measured parity errors of the CUDA paths against the golden fixtures (run on the GPU box):
    python tools/parity_report.py > gpurun_out/parity.txt"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "audio-analyzer-omega_b200"))
from omega4_b200 import _native as N
from omega4_b200.plan import AnalysisPlan, BASELINE_CONFIGS


def db_err(got, ref, floor_db=-60.0):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    peak = ref.max(axis=-1, keepdims=True)
    sig = ref >= np.maximum(peak * 10 ** (floor_db / 20), 1e-30)
    e = np.abs(20 * np.log10(np.maximum(got[sig], 1e-30)) - 20 * np.log10(ref[sig]))
    quiet = np.abs(got - ref)[~sig] / np.broadcast_to(peak, ref.shape)[~sig] if (~sig).any() else np.zeros(1)
    return e.max(), quiet.max()


g = np.load(os.path.join(ROOT, "tests", "golden", "multires_baseline.npz"))
p = AnalysisPlan(48000, BASELINE_CONFIGS, 512)
for name, fl in (("tensor-core blockdft (default)", 0), ("fp32 blockdft", N.FLAG_NO_TENSOR), ("full FFT", N.FLAG_NO_BLOCKDFT)):
    c = p.analyze_host(g["x"][None, :], want_meters=False, flags=fl)["combined"][0]
    e, q = db_err(c[16:], g["combined"][16:])
    e_lo, _ = db_err(c[16:, 1:26], g["combined"][16:, 1:26])
    print(f"{name:32s} max dB err (bins within 60 dB of the row max) {e:.2e}; bins 1..25 alone vs their own max {e_lo:.2e}; "
          f"quiet-bin abs err / row max {q:.2e}")
gm = np.load(os.path.join(ROOT, "tests", "golden", "meters_stream.npz"))
out = p.analyze_host(gm["x"][None, :], want_combined=False, want_series=True)
f = int(gm["first_hop"])
print(f"LUFS per-frame max err {np.abs(out['lufs_inst'][0, f:] - gm['lufs_inst']).max():.2e} LU; "
      f"true peak max err {np.abs(out['tp_db'][0, f:] - gm['tp_db']).max():.2e} dBTP; "
      f"meters (M,S,I,LRA,TP) max err {np.abs(out['meters'][0, f:] - gm['meters']).max(axis=0)}")
