"""GPU-backed data side of ``omega4.panels.spectrogram_waterfall.SpectrogramWaterfall`` (reference file
omega4/panels/spectrogram_waterfall.py) -- SURVEY.md section 8f rank 3.

``_setup_frequency_mapping`` (:55-69) is table construction and stays on the host; the per-frame
arithmetic of ``update`` / ``_normalize_spectrum`` (:71-121) -- slice, dB, row extrema, the 20-entry
P95 / P5 auto gain, normalise, clip -- runs in libomega4_cuda.so (``omega4_waterfall``).  The attribute
names are the ones the reference's draw code reads (``waterfall_data``, ``current_peak``,
``current_floor``, ``freq_indices``, ``display_freqs``, ``peak_history``); pygame drawing stays with the
reference panel.  No CPU fallback.
"""
from __future__ import annotations

from collections import deque
from typing import Optional

import numpy as np

from .. import _native as N


class SpectrogramWaterfall:
    def __init__(self, sample_rate: int = 48000, fft_size: int = 2048, device: int = 0):
        self.sample_rate = sample_rate
        self.fft_size = fft_size
        self.device = device
        self.waterfall_height = 200
        self.waterfall_data = deque(maxlen=self.waterfall_height)
        self.min_freq = 20
        self.max_freq = 20000
        self.freq_scale = "log"
        self.dynamic_range = 80
        self.color_scheme = "spectrum"
        self.auto_gain = True
        self.gain_adjustment = 0.0
        self.peak_history = deque(maxlen=100)
        self.current_peak = 0.0
        self.current_floor = -80.0
        self.freq_bins = None
        self.freq_indices = None
        self._setup_frequency_mapping()
        self.update_counter = 0
        self.update_interval = 1
        self._state = np.zeros((1, N.WATERFALL_STATE), np.float32)
        self._fresh = True

    def _setup_frequency_mapping(self):
        """spectrogram_waterfall.py:55-69."""
        nyquist = self.sample_rate / 2
        self.freq_bins = np.linspace(0, nyquist, self.fft_size // 2 + 1)
        min_idx = int(np.argmax(self.freq_bins >= self.min_freq))
        max_idx = int(np.argmax(self.freq_bins >= self.max_freq))
        if max_idx == 0:
            max_idx = len(self.freq_bins) - 1
        self.freq_indices = (min_idx, max_idx)
        self.display_freqs = self.freq_bins[min_idx:max_idx]

    # ------------------------------------------------------------------ per frame (application call pattern)
    def update(self, fft_data: np.ndarray, frequencies: Optional[np.ndarray] = None):
        """spectrogram_waterfall.py:71-105: one spectrum in, one normalised row appended to
        ``waterfall_data``."""
        self.update_counter += 1
        if self.update_counter % self.update_interval != 0:
            return
        if fft_data is None or len(fft_data) == 0:
            return
        rows = self.update_batch(np.asarray(fft_data)[None, :])
        return rows[0] if rows is not None else None

    # ------------------------------------------------------------------ batches
    def update_batch(self, spectra: np.ndarray, want_db: bool = False):
        """spectra [n_rows, len] in time order: the rows ``update`` would have appended one by one
        (float32 [n_rows, hi - lo]); ``want_db`` also returns the dB rows."""
        x = np.ascontiguousarray(spectra, dtype=np.float32)
        n_rows, ln = x.shape
        lo, hi = self.freq_indices
        if hi > ln or hi <= lo:
            raise N.Omega4CudaError(f"spectrum of {ln} bins does not cover the display slice [{lo}, {hi})")
        N.require_device()
        norm = np.empty((n_rows, hi - lo), np.float32)
        dbv = np.empty((n_rows, hi - lo), np.float32) if want_db else None
        stat = np.empty((n_rows, 4), np.float32)
        if not self.auto_gain:                              # the panel's attributes are the fixed range
            self._state[0, N.WATERFALL_STATE - 2:] = (self.current_peak, self.current_floor)
        rc = N.lib().omega4_waterfall(self.device, None, N.MEM_HOST, x.ctypes.data, 1, n_rows, ln, lo, hi, 0,
                                      1 if self.auto_gain else 0, float(self.gain_adjustment), self._state.ctypes.data,
                                      1 if (self._fresh and self.auto_gain) else 0, N.ptr(dbv), norm.ctypes.data,
                                      stat.ctypes.data)
        N.check(rc, "omega4_waterfall")
        self._fresh = False
        for k in range(n_rows):
            self.peak_history.append((stat[k, 0], stat[k, 1]))
            self.waterfall_data.append(norm[k])
        self.current_peak, self.current_floor = stat[-1, 2], stat[-1, 3]
        return (norm, dbv) if want_db else norm


def spectrogram_db(fft_data: np.ndarray, device: int = 0) -> np.ndarray:
    """SpectrogramPanel.update's conversion (omega4/plugins/panels/spectrogram.py:72):
    ``20 * log10(fft_data + 1e-10)`` over one spectrum or a batch of rows."""
    x = np.ascontiguousarray(fft_data, dtype=np.float32)
    one = x.ndim == 1
    if one:
        x = x[None, :]
    N.require_device()
    out = np.empty_like(x)
    rc = N.lib().omega4_waterfall(device, None, N.MEM_HOST, x.ctypes.data, 1, x.shape[0], x.shape[1], 0, x.shape[1], 1, 0, 0.0,
                                  None, 1, out.ctypes.data, None, None)
    N.check(rc, "omega4_waterfall")
    return out[0] if one else out
