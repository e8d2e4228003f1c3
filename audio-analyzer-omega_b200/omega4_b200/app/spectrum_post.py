"""Spectrum post-processing of the application frame loop on the GPU.

Mirrors the inline block of ``ProfessionalLiveAudioAnalyzer.process_audio_spectrum``
(omega4_main.py:992-1056) that turns the combined multi-resolution spectrum into the drawable
``band_values``: P98 normalisation x 0.8, ``apply_frequency_compensation`` (:855-926), optional
max normalisation, mel band mean -> sqrt -> clamp, per-band exponential smoothing.  The attribute
names (``freq_compensation_enabled``, ``normalization_enabled``, ``smoothing_enabled``,
``current_content_type``, ``vocal_suppression``, ``prev_band_values``, ``band_indices``) are the
application's own.  All arithmetic runs in libomega4_cuda (``omega4_bars_run``); no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from .. import _native as N
from .. import tables

SAMPLE_RATE = 48000          # omega4/config/config.py:7
FFT_SIZE_BASE = 2048         # omega4/config/config.py:12
BARS_DEFAULT = 512


class SpectrumPostProcessor:
    """``band_values, peak_values = post.process(spectrum)`` once per frame, or
    ``post.process_batch(combined[n_ch, n_hops, T])`` for a resident batch (torch CUDA tensors)."""

    def __init__(self, bars: int = BARS_DEFAULT, sample_rate: int = SAMPLE_RATE, fft_size_base: int = FFT_SIZE_BASE,
                 device: int = 0):
        self.bars = int(bars)
        self.sample_rate, self.fft_size_base, self.device = int(sample_rate), int(fft_size_base), int(device)
        self.freq_compensation_enabled = True      # omega4_main.py:158
        self.normalization_enabled = False         # :160
        self.smoothing_enabled = True              # :161
        self.current_content_type = "instrumental"  # :175
        self.vocal_suppression = 0.0               # :346
        self.freqs = np.fft.rfftfreq(self.fft_size_base, 1 / self.sample_rate)     # :168
        self.band_indices = tables.mel_band_indices(self.sample_rate, self.fft_size_base, self.bars)   # :171-178
        self.prev_band_values: Optional[np.ndarray] = None
        self._handle = None
        self._key = None
        self._state = None

    # ------------------------------------------------------------------ native object
    def _settings(self):
        return (self.bars, self.freq_compensation_enabled, self.normalization_enabled, self.smoothing_enabled,
                self.current_content_type, float(self.vocal_suppression))

    def _ensure(self):
        key = self._settings()
        if self._handle is not None and key == self._key:
            return self._handle
        self.close()
        lib = N.lib()
        N.require_device()
        bands = np.ascontiguousarray(self.band_indices, dtype=np.int32)
        d = N.BarsDesc()
        d.spectrum_len, d.n_bars = self.bars, len(bands)
        d.bands = bands.ctypes.data_as(C.POINTER(C.c_int))
        gain = smooth = None
        if self.freq_compensation_enabled:
            gain = tables.app_compensation_gains(self.freqs[:self.bars], self.current_content_type, self.vocal_suppression)
            gain = np.ascontiguousarray(gain, dtype=np.float32)
            d.gain = gain.ctypes.data_as(C.POINTER(C.c_float))
        if self.smoothing_enabled:
            smooth = np.ascontiguousarray(tables.app_smoothing_factors(self.band_indices, self.sample_rate,
                                                                       self.fft_size_base), dtype=np.float64)
            d.smooth = smooth.ctypes.data_as(C.POINTER(C.c_double))
        d.percentile, d.scale, d.normalize_max = 98.0, 0.8, int(self.normalization_enabled)
        h = lib.omega4_bars_create(C.byref(d), self.device)
        if not h:
            raise N.Omega4CudaError(f"omega4_bars_create failed: {N.last_error()}")
        self._handle, self._key = h, key
        self.n_valid = int(lib.omega4_bars_count(h))
        return h

    def close(self):
        if self._handle:
            N.lib().omega4_bars_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ per-frame (application call pattern)
    def process(self, spectrum) -> Tuple[np.ndarray, np.ndarray]:
        """One frame: returns (band_values, peak_values) like omega4_main.py:1035-1036 and keeps
        ``prev_band_values`` for the next frame (:1056)."""
        h = self._ensure()
        spec = np.ascontiguousarray(spectrum, dtype=np.float32).reshape(1, 1, -1)
        if spec.shape[-1] != self.bars:
            raise N.Omega4CudaError(f"spectrum length {spec.shape[-1]} != bars {self.bars}")
        if self._state is None or self._state.shape[1] != 1 + self.n_valid:
            self._state = np.zeros((1, 1 + self.n_valid), np.float32)
        if self.prev_band_values is None:
            self._state[0, 0] = 0.0
        else:
            self._state[0, 0] = 1.0
            self._state[0, 1:] = self.prev_band_values
        band = np.empty((1, 1, self.n_valid), np.float32)
        peak = np.empty_like(band)
        rc = N.lib().omega4_bars_run(h, None, N.MEM_HOST, spec.ctypes.data, 1, 1, self._state.ctypes.data, 0,
                                     band.ctypes.data, peak.ctypes.data)
        N.check(rc, "omega4_bars_run")
        self.prev_band_values = self._state[0, 1:].copy()
        return band[0, 0], peak[0, 0]

    # ------------------------------------------------------------------ batches
    def process_host(self, spectra: np.ndarray, state: Optional[np.ndarray] = None, want_peaks: bool = False):
        """spectra float32 [n_ch, n_hops, T] (host) -> band_values [n_ch, n_hops, n_valid] (+ peaks)."""
        h = self._ensure()
        x = np.ascontiguousarray(spectra, dtype=np.float32)
        n_ch, n_hops, t = x.shape
        if t != self.bars:
            raise N.Omega4CudaError(f"spectrum length {t} != bars {self.bars}")
        band = np.empty((n_ch, n_hops, self.n_valid), np.float32)
        peak = np.empty_like(band) if want_peaks else None
        rc = N.lib().omega4_bars_run(h, None, N.MEM_HOST, x.ctypes.data, n_ch, n_hops, N.ptr(state),
                                     0 if state is not None else 1, band.ctypes.data, N.ptr(peak))
        N.check(rc, "omega4_bars_run")
        return (band, peak) if want_peaks else band

    def process_device(self, spectra, band_values, state=None, fresh: bool = True, peak_values=None, stream=None):
        """Device tensors (torch, float32, contiguous): spectra [n_ch, n_hops, T] -> band_values
        [n_ch, n_hops, n_valid]; asynchronous on ``stream`` (default: torch's current stream)."""
        import torch
        h = self._ensure()
        n_ch, n_hops, t = spectra.shape
        if t != self.bars or tuple(band_values.shape) != (n_ch, n_hops, self.n_valid):
            raise N.Omega4CudaError("bad tensor shapes for process_device")
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        rc = N.lib().omega4_bars_run(h, s, N.MEM_DEVICE, spectra.data_ptr(), n_ch, n_hops, N.ptr(state),
                                     1 if (fresh or state is None) else 0, band_values.data_ptr(), N.ptr(peak_values))
        N.check(rc, "omega4_bars_run")
        return band_values
