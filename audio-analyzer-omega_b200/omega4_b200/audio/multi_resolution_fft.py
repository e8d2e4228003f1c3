"""GPU-backed drop-in for ``omega4.audio.multi_resolution_fft`` (reference file
omega4/audio/multi_resolution_fft.py).  Same public names, argument meaning and return types:

    MultiResolutionFFT(sample_rate=48000, max_freq=20000)
        .configs, .windows, .buffers, .freq_arrays
        .process_audio_chunk(audio_chunk, apply_weighting=True) -> Dict[int, FFTResult]
        .combine_results_optimized(results, target_bins=1024) -> (float32[T], float64[T])
        .get_frequency_arrays(), .reset_all_buffers(), .get_processing_stats(),
        .get_buffer_status(), .cleanup()

Window x FFT x |.| x psychoacoustic weights and the combine interpolation run in
libomega4_cuda.so (multires_kernel / combine_kernel); this module only keeps the ring-buffer
bookkeeping of ``CircularBuffer`` (:52-133) on the host.  Differences from the reference, both
deliberate: (1) a missing CUDA library / device raises instead of silently computing on the CPU
(the product has no CPU path); (2) nothing else -- including the ``WindowType.HANN`` ->
rectangular-window quirk (:179-193) and the per-call ``magnitude.copy()`` ownership (:283).
"""
from __future__ import annotations

import ctypes as C
import logging
import threading
import time
from dataclasses import dataclass
from enum import Enum
from typing import Dict, List, NamedTuple, Optional, Tuple

import numpy as np

from .. import _native as N
from .. import tables
from ..plan import AnalysisPlan, ResSpec

logger = logging.getLogger(__name__)


class WindowType(Enum):
    """Window function types (:19-24)."""
    BLACKMAN = "blackman"
    HANN = "hann"
    HAMMING = "hamming"
    BLACKMAN_HARRIS = "blackman_harris"


@dataclass
class FFTConfig:
    """Configuration for a single FFT resolution (:26-44)."""
    freq_range: Tuple[float, float]
    fft_size: int
    hop_size: int
    weight: float
    window_type: WindowType = WindowType.BLACKMAN

    def __post_init__(self):
        if self.freq_range[0] >= self.freq_range[1]:
            raise ValueError(f"Invalid frequency range: {self.freq_range}")
        if self.fft_size <= 0 or (self.fft_size & (self.fft_size - 1)) != 0:
            raise ValueError(f"FFT size must be power of 2: {self.fft_size}")
        if self.hop_size <= 0:
            raise ValueError(f"Hop size must be positive: {self.hop_size}")
        if self.weight <= 0:
            raise ValueError(f"Weight must be positive: {self.weight}")


class FFTResult(NamedTuple):
    """Result from FFT processing (:46-50)."""
    magnitude: np.ndarray
    frequencies: np.ndarray
    config_index: int


class CircularBuffer:
    """Thread-safe float32 ring (:52-133): oversize writes keep the last ``size`` samples,
    ``read_latest`` returns None until ``length`` samples have been written."""

    def __init__(self, size: int, dtype=np.float32):
        if size <= 0:
            raise ValueError("Buffer size must be positive")
        self.size = size
        self.buffer = np.zeros(size, dtype=dtype)
        self.write_pos = 0
        self.samples_written = 0
        self._lock = threading.Lock()

    def write(self, data) -> bool:
        if data is None or len(data) == 0:
            return False
        with self._lock:
            n = len(data)
            if n >= self.size:
                self.buffer[:] = data[-self.size:]
                self.write_pos = 0
                self.samples_written = self.size
            else:
                end = self.write_pos + n
                if end <= self.size:
                    self.buffer[self.write_pos:end] = data
                else:
                    k = self.size - self.write_pos
                    self.buffer[self.write_pos:] = data[:k]
                    self.buffer[:n - k] = data[k:]
                self.write_pos = end % self.size
                self.samples_written = min(self.samples_written + n, self.size)
        return True

    def read_latest(self, length: int) -> Optional[np.ndarray]:
        if length <= 0 or length > self.size:
            return None
        with self._lock:
            if self.samples_written < length:
                return None
            start = self.write_pos - length                     # same samples as the reference's modular index, two slices
            if start >= 0:
                return self.buffer[start:self.write_pos].copy()
            return np.concatenate((self.buffer[start:], self.buffer[:self.write_pos]))

    def reset(self):
        with self._lock:
            self.buffer.fill(0)
            self.write_pos = 0
            self.samples_written = 0


class MultiResolutionFFT:
    """Multi-resolution FFT analysis, GPU backed (:135-458)."""

    def __init__(self, sample_rate: int = 48000, max_freq: float = 20000, device: int = 0):
        if sample_rate <= 0:
            raise ValueError("Sample rate must be positive")
        if max_freq <= 0 or max_freq > sample_rate / 2:
            raise ValueError("Max frequency must be positive and <= Nyquist")
        self.sample_rate = sample_rate
        self.nyquist = sample_rate / 2
        self.max_freq = min(max_freq, self.nyquist)
        self.device = device
        self.configs = [
            FFTConfig((20, 200), 4096, 1024, 1.5),
            FFTConfig((200, 1000), 2048, 512, 1.2),
            FFTConfig((1000, 5000), 1024, 256, 1.0),
            FFTConfig((5000, 20000), 1024, 256, 1.5),
        ]
        self._plans: Dict[Tuple[bool, int], AnalysisPlan] = {}
        self._combine_bins = 1024                           # combine_results_optimized's default target_bins (:335)
        self._last = None
        self._setup_windows()
        self._setup_buffers()
        self._setup_frequency_arrays()
        self._setup_working_arrays()
        self.processing_stats = {"total_calls": 0, "total_time": 0.0, "error_count": 0}
        logger.info(f"MultiResolutionFFT initialized: {sample_rate}Hz, {len(self.configs)} resolutions")

    # -- setup (callers overwrite .configs and re-run these, as the reference's users do) ------
    def _invalidate(self):
        for p in getattr(self, "_plans", {}).values():
            p.close()
        self._plans = {}
        self._last = None

    def _setup_windows(self):
        self.windows = {i: tables.multires_window(c.window_type.value if isinstance(c.window_type, WindowType)
                                                  else str(c.window_type), c.fft_size)
                        for i, c in enumerate(self.configs)}
        self._invalidate()

    def _setup_buffers(self):
        self.buffers = {i: CircularBuffer(max(c.fft_size * 2, c.fft_size + c.hop_size))
                        for i, c in enumerate(self.configs)}

    def _setup_frequency_arrays(self):
        self.freq_arrays = {i: np.fft.rfftfreq(c.fft_size, 1 / self.sample_rate) for i, c in enumerate(self.configs)}
        self._invalidate()

    def _setup_working_arrays(self):
        self.working_arrays = {i: {"audio_data": np.zeros(c.fft_size, dtype=np.float32)}
                               for i, c in enumerate(self.configs)}

    def _plan(self, apply_weighting: bool, target_bins: int) -> AnalysisPlan:
        key = (bool(apply_weighting), int(target_bins))
        p = self._plans.get(key)
        if p is None:
            specs = [ResSpec(tuple(c.freq_range), c.fft_size, c.hop_size, c.weight,
                             c.window_type.value if isinstance(c.window_type, WindowType) else str(c.window_type))
                     for c in self.configs]
            p = AnalysisPlan(self.sample_rate, specs, target_bins, self.max_freq, hop=512,
                             apply_weighting=apply_weighting, device=self.device,
                             windows=[self.windows[i] for i in range(len(self.configs))])
            self._plans[key] = p
        return p

    # -- hot path ------------------------------------------------------------------------------
    def process_audio_chunk(self, audio_chunk: np.ndarray, apply_weighting: bool = True) -> Dict[int, FFTResult]:
        if audio_chunk is None or len(audio_chunk) == 0:
            logger.warning("Empty audio chunk received")
            return {}
        start = time.perf_counter()
        ready: List[int] = []
        frames: List[np.ndarray] = []
        for i, c in enumerate(self.configs):
            if not self.buffers[i].write(audio_chunk):
                continue
            a = self.buffers[i].read_latest(c.fft_size)
            if a is None:
                continue
            ready.append(i)
            frames.append(a)
        results: Dict[int, FFTResult] = {}
        if ready:
            # ONE host round trip for every ready resolution: the frames go up together, the magnitudes (and the
            # combination over exactly these resolutions, for the combine_results_optimized call that follows in
            # the application, omega4_main.py:707-714) come back together (omega4_stream_hop)
            plan = self._plan(apply_weighting, self._combine_bins)
            n_res = len(self.configs)
            fptr = (C.c_void_p * n_res)()
            mptr = (C.c_void_p * n_res)()
            mags = {}
            keep = []
            for i, a in zip(ready, frames):
                a = np.ascontiguousarray(a, dtype=np.float32)
                keep.append(a)
                mags[i] = np.empty(len(a) // 2 + 1, np.float32)
                fptr[i] = a.ctypes.data
                mptr[i] = mags[i].ctypes.data
            comb = np.empty(self._combine_bins, np.float32)
            rc = N.lib().omega4_stream_hop(plan.handle, fptr, mptr, comb.ctypes.data)
            N.check(rc, "omega4_stream_hop")
            for i in ready:
                results[i] = FFTResult(magnitude=mags[i], frequencies=self.freq_arrays[i], config_index=i)
            # remembered so that combine_results_optimized(results) can hand the combination back without another
            # round trip -- only if the caller passes these very magnitudes, unmodified (compared value by value)
            self._last = (bool(apply_weighting), self._combine_bins, {i: mags[i].copy() for i in ready}, comb)
        self.processing_stats["total_calls"] += 1
        self.processing_stats["total_time"] += time.perf_counter() - start
        return results

    def combine_results_optimized(self, results: Dict[int, FFTResult], target_bins: int = 1024):
        if not results:
            logger.warning("No FFT results to combine")
            return np.zeros(target_bins), np.linspace(0, self.max_freq, target_bins)
        last = self._last
        if last is not None and last[1] == target_bins and len(results) == len(last[2]) and \
                all(r.config_index in last[2] and np.array_equal(r.magnitude, last[2][r.config_index]) for r in results.values()):
            return last[3].copy(), np.linspace(0, self.max_freq, target_bins)
        self._combine_bins = int(target_bins)               # the next process_audio_chunk combines for this length
        plan = self._plan(True, target_bins)
        mags = [None] * len(self.configs)
        for r in results.values():
            mags[r.config_index] = np.asarray(r.magnitude, dtype=np.float32)
        combined = plan.combine_host(mags, 1)[0]
        self._last = (True, int(target_bins), {r.config_index: np.array(r.magnitude, dtype=np.float32, copy=True) for r in results.values()},
                      combined.copy())
        return combined, np.linspace(0, self.max_freq, target_bins)

    # -- bookkeeping API of the reference ------------------------------------------------------
    def get_frequency_arrays(self) -> Dict[int, np.ndarray]:
        return self.freq_arrays.copy()

    def reset_all_buffers(self):
        for b in self.buffers.values():
            b.reset()
        logger.info("All buffers reset")

    def get_processing_stats(self) -> Dict[str, float]:
        s = self.processing_stats.copy()
        if s["total_calls"] > 0:
            s["avg_time_ms"] = (s["total_time"] / s["total_calls"]) * 1000
            s["error_rate"] = s["error_count"] / s["total_calls"]
        else:
            s["avg_time_ms"] = 0.0
            s["error_rate"] = 0.0
        return s

    def get_buffer_status(self) -> Dict[int, Dict[str, int]]:
        return {i: {"size": b.size, "write_pos": b.write_pos, "samples_written": b.samples_written,
                    "utilization_pct": int((b.samples_written / b.size) * 100)} for i, b in self.buffers.items()}

    def cleanup(self):
        self.reset_all_buffers()
        self._invalidate()
        logger.info("MultiResolutionFFT cleanup completed")


def create_default_multi_fft(sample_rate: int = 48000) -> MultiResolutionFFT:
    return MultiResolutionFFT(sample_rate=sample_rate)


def benchmark_multi_fft(sample_rate: int = 48000, chunk_size: int = 512, num_iterations: int = 1000) -> Dict[str, float]:
    """Same driver as the reference's in-module benchmark (:467-494)."""
    proc = create_default_multi_fft(sample_rate)
    audio = np.random.random(chunk_size).astype(np.float32)
    for _ in range(10):
        proc.process_audio_chunk(audio)
    t0 = time.perf_counter()
    for _ in range(num_iterations):
        res = proc.process_audio_chunk(audio)
        if res:
            proc.combine_results_optimized(res)
    total = time.perf_counter() - t0
    return {"total_time_s": total, "avg_time_ms": (total / num_iterations) * 1000,
            "iterations_per_second": num_iterations / total, "stats": proc.get_processing_stats()}
