"""GPU-backed data side of ``omega4.panels.bass_zoom.BassZoomPanel`` (reference file
omega4/panels/bass_zoom.py) -- SURVEY.md section 8f rank 3.

``setup_bass_mapping`` (:50-97) is table construction and stays on the host, bit for bit; the
per-frame arithmetic of ``_process_bass_detail_internal`` (:141-214) -- Hann window, zero padding to
8192, rFFT, magnitude (``omega4_rfft_batch``), per-bar mean, frequency compensation, dynamic scaling,
compression, attack/release smoothing, clamp (``omega4_bass_bars``) -- runs in libomega4_cuda.so.
The wall-clock peak hold (:206-211) is three lines of host bookkeeping on ``time.time()`` exactly as
in the reference; the worker thread / queue of the reference (:32-38, :99-115) is not needed because
the GPU call is the asynchronous part.  pygame drawing stays with the reference panel.
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import numpy as np

from .. import _native as N

BASS_FFT_SIZE = 8192


class BassZoomPanel:
    def __init__(self, sample_rate: int = 48000, device: int = 0):
        self.sample_rate = sample_rate
        self.device = device
        self.bass_detail_bars = 64
        self.bass_bar_values = np.zeros(self.bass_detail_bars, dtype=np.float32)
        self.bass_peak_values = np.zeros(self.bass_detail_bars, dtype=np.float32)
        self.bass_peak_timestamps = np.zeros(self.bass_detail_bars, dtype=np.float64)
        self.bass_freq_ranges = []
        self.bass_bin_mapping = []
        self.drum_info: Dict = {}
        self.setup_bass_mapping()

    def setup_bass_mapping(self):
        """bass_zoom.py:50-97, verbatim semantics."""
        bass_freqs = np.fft.rfftfreq(BASS_FFT_SIZE, 1 / self.sample_rate)
        valid_bins = [i for i, f in enumerate(bass_freqs) if 20 <= f <= 200]
        if len(valid_bins) == 0:
            self.bass_freq_ranges = [(20, 200)]
            self.bass_bin_mapping = [[]]
            self.bass_detail_bars = 1
        else:
            self.bass_freq_ranges, self.bass_bin_mapping = [], []
            bins_per_bar = max(1, len(valid_bins) // 31)
            for i in range(0, len(valid_bins), bins_per_bar):
                group = valid_bins[i:min(i + bins_per_bar, len(valid_bins))]
                if len(group) > 0:
                    f_start, f_end = bass_freqs[group[0]], bass_freqs[group[-1]]
                    if len(self.bass_freq_ranges) > 0:
                        f_start = max(f_start, self.bass_freq_ranges[-1][1])
                    self.bass_freq_ranges.append((f_start, f_end))
                    self.bass_bin_mapping.append(group)
            self.bass_detail_bars = len(self.bass_freq_ranges)
        self.bass_bar_values = np.zeros(self.bass_detail_bars, dtype=np.float32)
        self.bass_peak_values = np.zeros(self.bass_detail_bars, dtype=np.float32)
        self.bass_peak_timestamps = np.zeros(self.bass_detail_bars, dtype=np.float64)
        # tables for the kernel: contiguous bin groups (first, count) and the compensation of :167-174
        self._bar_bins = np.array([[g[0], len(g)] if len(g) else [0, 0] for g in self.bass_bin_mapping], dtype=np.int32)
        comp = []
        for i in range(self.bass_detail_bars):
            c = (self.bass_freq_ranges[i][0] + self.bass_freq_ranges[i][1]) / 2
            comp.append(0.3 if c < 60 else 0.6 if c < 100 else 1.0 if c < 150 else 0.8)
        self._comp = np.array(comp, dtype=np.float32)

    # ------------------------------------------------------------------ GPU arithmetic
    def bar_values_batch(self, frames: np.ndarray, state: Optional[np.ndarray] = None) -> np.ndarray:
        """frames [n_ch, n_frames, L <= 8192] (or [n_frames, L]) -> bass bar values [n_ch, n_frames, bars];
        ``state`` float32 [n_ch, bars] carries the previous bars (None: zeros)."""
        f = np.asarray(frames)
        if f.ndim == 2:
            f = f[None]
        n_ch, n_frames, ln = f.shape
        ln_eff = min(ln, BASS_FFT_SIZE)
        padded = np.zeros((n_ch * n_frames, BASS_FFT_SIZE), np.float32)        # zero padding: data movement only
        padded[:, :ln_eff] = f.reshape(n_ch * n_frames, ln)[:, :ln_eff]
        window = np.zeros(BASS_FFT_SIZE, np.float32)
        window[:ln_eff] = np.hanning(ln_eff)                                   # table; the multiply is in the kernel
        mag = np.empty((n_ch * n_frames, BASS_FFT_SIZE // 2 + 1), np.float32)
        N.require_device()
        rc = N.lib().omega4_rfft_batch(self.device, None, N.MEM_HOST, N.ptr(padded), n_ch * n_frames, BASS_FFT_SIZE,
                                       N.ptr(window), N.ptr(mag), None)
        N.check(rc, "omega4_rfft_batch")
        out = np.empty((n_ch, n_frames, self.bass_detail_bars), np.float32)
        rc = N.lib().omega4_bass_bars(self.device, None, N.MEM_HOST, N.ptr(mag), n_ch, n_frames, mag.shape[1],
                                      N.ptr(self._bar_bins), N.ptr(self._comp), self.bass_detail_bars, N.ptr(state), N.ptr(out))
        N.check(rc, "omega4_bass_bars")
        return out

    def _process_bass_detail_internal(self, audio_data: np.ndarray):
        state = self.bass_bar_values.astype(np.float32)[None, :].copy()
        bars = self.bar_values_batch(np.asarray(audio_data)[None, None, :], state)[0, 0]
        peak_values = self.bass_peak_values.copy()
        peak_timestamps = self.bass_peak_timestamps.copy()
        now = time.time()
        for i, group in enumerate(self.bass_bin_mapping):                      # peak hold, :206-211
            if len(group) == 0:
                continue
            if bars[i] > peak_values[i]:
                peak_values[i] = bars[i]
                peak_timestamps[i] = now
            elif now - peak_timestamps[i] > 3.0:
                peak_values[i] *= 0.95
        return {"bar_values": bars, "peak_values": peak_values, "peak_timestamps": peak_timestamps}

    def update(self, audio_data: np.ndarray, drum_info: Dict = None):
        res = self._process_bass_detail_internal(audio_data)
        self.bass_bar_values = res["bar_values"]
        self.bass_peak_values = res["peak_values"]
        self.bass_peak_timestamps = res["peak_timestamps"]
        self.drum_info = drum_info if drum_info is not None else {}

    def get_results(self) -> Dict:
        return {"bar_values": self.bass_bar_values, "peak_values": self.bass_peak_values}
