"""GPU-backed drop-in for ``omega4.optimization.freq_mapper`` (reference file
omega4/optimization/freq_mapper.py): ``PrecomputedFrequencyMapper`` / ``FrequencyMapping``.

The mel band ``(start, end)`` tables are host integers computed with the reference's own
arithmetic (bit exact, :83-124); ``map_spectrum_to_bars`` (:165-196) runs as ``band_map_kernel``.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from .. import tables
from ..plan import band_map_host

logger = logging.getLogger(__name__)


@dataclass
class FrequencyMapping:
    band_indices: List[Tuple[int, int]]
    freq_to_bin: np.ndarray
    bin_to_freq: np.ndarray
    mel_scale_factors: np.ndarray
    compensation_curve: np.ndarray
    frequency_points: np.ndarray


class PrecomputedFrequencyMapper:
    def __init__(self, sample_rate: int, fft_size: int, num_bars: int, device: int = 0):
        self.sample_rate = sample_rate
        self.fft_size = fft_size
        self.num_bars = num_bars
        self.device = device
        self.freq_bin_width = sample_rate / fft_size
        self.mapping = self._precompute_all()
        self.interp_cache = {}
        logger.info(f"Pre-computed frequency mappings for {num_bars} bars, FFT size {fft_size}, sample rate {sample_rate}")

    def _precompute_all(self) -> FrequencyMapping:
        bands = tables.mel_band_indices(self.sample_rate, self.fft_size, self.num_bars)
        freqs = np.arange(self.fft_size // 2 + 1) * self.freq_bin_width
        points = np.zeros(self.num_bars)
        for i, (s, e) in enumerate(bands):
            if i < self.num_bars:
                points[i] = ((s + e) // 2) * self.freq_bin_width
        return FrequencyMapping(band_indices=bands, freq_to_bin=freqs, bin_to_freq=freqs.copy(),
                                mel_scale_factors=tables.mel_scale_factors(freqs),
                                compensation_curve=tables.compensation_curve(freqs), frequency_points=points)

    def map_spectrum_to_bars(self, spectrum: np.ndarray, apply_compensation: bool = True) -> np.ndarray:
        spectrum = np.asarray(spectrum)
        comp = None
        if apply_compensation and len(spectrum) == len(self.mapping.compensation_curve):
            comp = self.mapping.compensation_curve
        return band_map_host(spectrum, self.mapping.band_indices, comp, db=False, device=self.device)[0]

    def get_frequency_for_bar(self, bar_index: int) -> float:
        if 0 <= bar_index < self.num_bars:
            return self.mapping.frequency_points[bar_index]
        return 0.0

    def get_bar_for_frequency(self, frequency: float) -> int:
        idx = np.searchsorted(self.mapping.frequency_points, frequency)
        return max(0, min(idx, self.num_bars - 1))
