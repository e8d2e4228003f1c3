"""TEST / BASELINE INFRASTRUCTURE ONLY -- per-hop CPU port with the reference's cost profile.

``oracle_np`` restates the algorithm in batched numpy (fast enough for tests).  The CPU
*baseline* must instead cost what the reference costs: one Python call per hop per resolution,
the same library primitives (``np.fft.rfft``, ``scipy.signal.filtfilt``, ``scipy.signal.resample``,
``np.percentile``) at the same granularity.  This module is that port -- a rewrite of the call
structure of

    MultiResolutionFFT.process_audio_chunk / combine_results_optimized   (omega4/audio/multi_resolution_fft.py:228-408)
    ProfessionalMetering.calculate_lufs / calculate_true_peak             (omega4/panels/professional_meters.py:231-299)

driven on the shared hop schedule.  ``tests/test_oracle_golden.py::test_ref_port_matches_oracle``
pins it to ``oracle_np`` (and through it to the reference goldens).  Used only by ``bench.py``'s
``cpu_baseline`` leg and ``--impl reference`` arm (the reference itself is Python and does not
exist on the GPU box).
"""
from __future__ import annotations

import os
import time
from collections import deque

import numpy as np

from . import oracle_np as O

try:                                   # the reference's own dependency; present in the image
    from scipy import signal as _sps
except Exception:                      # pragma: no cover
    _sps = None


class PortMultiRes(O.OracleMultiResFFT):
    """Per-chunk port; rebuilds the psychoacoustic weight vector on EVERY call like the reference
    does (multi_resolution_fft.py:309-326) -- that waste is part of the baseline's cost."""

    def bin_weights(self, i):
        c = self.configs[i]
        return O.psycho_weights(self.freq_arrays[i], c.freq_range, c.weight)


class PortMetering:
    """calculate_lufs with scipy's filtfilt/resample per frame (professional_meters.py:129-299)."""

    def __init__(self, sample_rate=48000):
        if _sps is None:
            raise RuntimeError("scipy is required for the CPU baseline port")
        self.c = O.k_weighting_coeffs(sample_rate)
        self.mom, self.short = deque(maxlen=24), deque(maxlen=180)
        self.integ, self.peaks = deque(maxlen=3600), deque(maxlen=60)
        self.cur = {"momentary": -100.0, "short_term": -100.0, "integrated": -100.0, "range": 0.0, "true_peak": -100.0}

    def calculate_lufs(self, x):
        if np.sqrt(np.mean(x ** 2)) < 1e-6:
            w = np.zeros_like(x)
        else:
            f = _sps.filtfilt(self.c["hp_b"], self.c["hp_a"], x)
            s = _sps.filtfilt(self.c["shelf_b"], self.c["shelf_a"], f)
            w = f + (s - f) * 0.3
        ms = np.mean(w ** 2)
        li = -0.691 + 10 * np.log10(ms) if ms > 1e-10 else -100.0
        self.mom.append(li); self.short.append(li); self.integ.append(li)
        self.cur["momentary"] = np.mean(self.mom)
        self.cur["short_term"] = np.mean(self.short)
        gated = [v for v in self.integ if v > -70.0]
        if gated:
            self.cur["integrated"] = np.mean(gated)
            self.cur["range"] = np.percentile(gated, 95) - np.percentile(gated, 10)
        else:
            self.cur["integrated"] = -100.0
            self.cur["range"] = 0.0
        peak = np.max(np.abs(_sps.resample(x, len(x) * 4)))
        self.peaks.append(-100.0 if peak < 1e-10 else 20 * np.log10(peak))
        self.cur["true_peak"] = max(self.peaks)
        return self.cur


def run_channel(x, sample_rate=48000, configs=O.BASELINE_CONFIGS, hop=512, target_bins=512, window=2048):
    """One channel, hop by hop, exactly the work the reference's CPU path does per hop."""
    mr = PortMultiRes(sample_rate, 20000, configs)
    met = PortMetering(sample_rate)
    hann = np.hanning(window)
    n_hops = len(x) // hop
    comb = np.zeros((n_hops, target_bins), np.float32)
    meters = np.zeros((n_hops, 5))
    for k in range(n_hops):
        res = mr.process_audio_chunk(x[k * hop:(k + 1) * hop], True)
        if res:
            comb[k] = mr.combine(res, target_bins)[0]
        e = (k + 1) * hop
        if e >= window:
            met.calculate_lufs(x[e - window:e] * hann)
        meters[k] = [met.cur[key] for key in O.METER_KEYS]
    return comb, meters


def _work(args):
    stream, n_channels, n_samples, sample_rate, configs = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "audio-analyzer-omega_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from omega4_b200.batch.synth import synth_channel      # the seeded generator only
    acc = 0.0
    for c in range(n_channels):
        x = synth_channel(stream, c, n_samples, sample_rate)
        t0 = time.perf_counter()
        comb, meters = run_channel(x, sample_rate, configs)
        acc += time.perf_counter() - t0
    return acc, float(meters[-1, 0])


def time_cpu_path(n_streams, n_channels=2, seconds=4.0, sample_rate=48000, configs=O.BASELINE_CONFIGS,
                  processes=None):
    """Stream-seconds analysed per wall-second by the CPU port on ``processes`` host cores
    (multiprocessing, OMP_NUM_THREADS=1, streams round-robin -- BASELINE.md section 3)."""
    import multiprocessing as mp
    procs = processes or os.cpu_count() or 1
    n_samples = int(seconds * sample_rate) // 512 * 512
    jobs = [(s, n_channels, n_samples, sample_rate, configs) for s in range(n_streams)]
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_work, [(0, 1, 16 * 512, sample_rate, configs)] * procs)   # warm the workers (imports, FFT plans)
        t0 = time.perf_counter()
        res = pool.map(_work, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    stream_seconds = n_streams * n_samples / sample_rate
    return {"value": stream_seconds / wall, "wall_s": wall, "cores": procs, "stream_seconds": stream_seconds,
            "cpu_s_per_stream_second": sum(r[0] for r in res) / stream_seconds}


if __name__ == "__main__":                       # `python -m oracle.ref_port --streams 8 --seconds 4`
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--processes", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=1)
    a = ap.parse_args()
    procs = a.processes or os.cpu_count() or 1
    out = [time_cpu_path(a.streams or procs, a.channels, a.seconds, processes=procs) for _ in range(a.repeat)]
    print(json.dumps(out))
