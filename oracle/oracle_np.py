"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the OMEGA-4 per-frame analysis path.

CPU oracle for the hot path named by BASELINE.json (multi-resolution FFT + combine +
professional meters).  It restates, in plain numpy, what the reference's Python does; every
function cites the reference file:line it follows (paths relative to /root/reference).

Third-party arithmetic the reference delegates to (not vendored in /root/reference;
requirements.txt:2-4 gives lower bounds only -- numpy>=1.21, scipy>=1.7; the container
that produced the golden vectors has numpy 2.3.5 / scipy 1.18.1):

* ``numpy.fft.rfft/irfft`` (pocketfft)     -- used here as the FFT primitive as well.
* ``numpy.blackman/hamming/hanning``       -- restated (``np_blackman`` ...).
* ``scipy.signal.butter(2, wn, 'high')``   -- restated in closed form (``butter2_highpass``).
* ``scipy.signal.lfilter_zi/lfilter``      -- restated (``lfilter_zi2``, ``lfilter_tdf2``).
* ``scipy.signal.filtfilt`` (padtype='odd', padlen=9, method='pad') -- restated (``filtfilt2``).
* ``scipy.signal.resample`` (FFT method, real input, even length)   -- restated (``resample_fft``).
* ``numpy.interp`` / ``numpy.percentile``  -- used as primitives.

PARITY IS PINNED: the reference has no golden vectors of its own, so
``oracle/gen_golden.py`` imports the UNMODIFIED reference classes in the build container and
freezes their outputs on seeded inputs into ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those fixtures.

Nothing here is imported by the product.
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# constants the path uses (omega4/config/config.py:7-12 -- the package, not omega4/config.py)
# --------------------------------------------------------------------------------------
SAMPLE_RATE = 48000
CHUNK_SIZE = 512
BARS_DEFAULT = 512
MAX_FREQ = 20000
FFT_SIZE_BASE = 2048

#: reference default resolutions, multi_resolution_fft.py:149-154: (freq_range, fft_size, hop, weight)
DEFAULT_CONFIGS = (
    ((20, 200), 4096, 1024, 1.5),
    ((200, 1000), 2048, 512, 1.2),
    ((1000, 5000), 1024, 256, 1.0),
    ((5000, 20000), 1024, 256, 1.5),
)
#: BASELINE.json sizes substituted in order ("Resolution <-> range mapping")
BASELINE_CONFIGS = (
    ((20, 200), 8192, 1024, 1.5),
    ((200, 1000), 4096, 512, 1.2),
    ((1000, 5000), 2048, 256, 1.0),
    ((5000, 20000), 1024, 256, 1.5),
)
#: BASELINE config 5 (96 kHz, six resolutions up to 32768): the reference defines no such list; ranges
#: chosen per SURVEY.md section 7 and fed to the unmodified reference for the golden vectors
CONFIG5_96K = (
    ((20, 60), 32768, 1024, 1.5),
    ((60, 200), 16384, 1024, 1.5),
    ((200, 1000), 8192, 512, 1.2),
    ((1000, 5000), 4096, 256, 1.0),
    ((5000, 12000), 2048, 256, 1.2),
    ((12000, 20000), 1024, 256, 1.5),
)


# --------------------------------------------------------------------------------------
# window functions (numpy's published formulas; all are the symmetric "M-1" forms)
# --------------------------------------------------------------------------------------
def _sym_n(m: int) -> np.ndarray:
    return np.arange(1 - m, m, 2)


def np_blackman(m: int) -> np.ndarray:
    """numpy.blackman: 0.42 + 0.5 cos(pi n/(M-1)) + 0.08 cos(2 pi n/(M-1)), n = 1-M, 3-M, ..., M-1."""
    if m < 1:
        return np.array([], dtype=np.float64)
    if m == 1:
        return np.ones(1, dtype=np.float64)
    n = _sym_n(m)
    return 0.42 + 0.5 * np.cos(np.pi * n / (m - 1)) + 0.08 * np.cos(2.0 * np.pi * n / (m - 1))


def np_hamming(m: int) -> np.ndarray:
    """numpy.hamming: 0.54 + 0.46 cos(pi n/(M-1))."""
    if m < 1:
        return np.array([], dtype=np.float64)
    if m == 1:
        return np.ones(1, dtype=np.float64)
    n = _sym_n(m)
    return 0.54 + 0.46 * np.cos(np.pi * n / (m - 1))


def np_hanning(m: int) -> np.ndarray:
    """numpy.hanning: 0.5 + 0.5 cos(pi n/(M-1))."""
    if m < 1:
        return np.array([], dtype=np.float64)
    if m == 1:
        return np.ones(1, dtype=np.float64)
    n = _sym_n(m)
    return 0.5 + 0.5 * np.cos(np.pi * n / (m - 1))


def multires_window(window_type: str, n: int) -> np.ndarray:
    """Window table of MultiResolutionFFT._setup_windows (multi_resolution_fft.py:171-193).

    BLACKMAN -> np.blackman; HAMMING -> np.hamming; BLACKMAN_HARRIS -> np.blackman (:183-184);
    HANN calls ``np.hann`` which does not exist -> AttributeError -> the except branch
    installs a rectangular window (:190-193).  Result is ``.astype(float32)`` (:188).
    """
    wt = window_type.lower()
    if wt in ("blackman", "blackman_harris"):
        w = np_blackman(n)
    elif wt == "hamming":
        w = np_hamming(n)
    elif wt == "hann":
        return np.ones(n, dtype=np.float32)
    else:
        w = np_blackman(n)
    return w.astype(np.float32)


def batched_window(window_type: str, n: int) -> np.ndarray:
    """BatchedFFTProcessor._get_window (batched_fft_processor.py:101-117): hann -> np.hanning,
    hamming, blackman, anything else -> ones; all ``.astype(float32)``."""
    if window_type == "hann":
        return np_hanning(n).astype(np.float32)
    if window_type == "hamming":
        return np_hamming(n).astype(np.float32)
    if window_type == "blackman":
        return np_blackman(n).astype(np.float32)
    return np.ones(n, dtype=np.float32)


def gpufft_window(window_type: str, n: int) -> np.ndarray:
    """GPUAcceleratedFFT window choice (gpu_accelerated_fft.py:115-125): hann, hamming,
    anything else -> blackman; float32."""
    if window_type == "hann":
        return np_hanning(n).astype(np.float32)
    if window_type == "hamming":
        return np_hamming(n).astype(np.float32)
    return np_blackman(n).astype(np.float32)


# --------------------------------------------------------------------------------------
# multi-resolution FFT  (omega4/audio/multi_resolution_fft.py)
# --------------------------------------------------------------------------------------
@dataclass
class OracleFFTConfig:
    """FFTConfig (multi_resolution_fft.py:26-44)."""
    freq_range: Tuple[float, float]
    fft_size: int
    hop_size: int
    weight: float
    window_type: str = "blackman"

    def __post_init__(self):
        if self.freq_range[0] >= self.freq_range[1]:
            raise ValueError(f"Invalid frequency range: {self.freq_range}")
        if self.fft_size <= 0 or (self.fft_size & (self.fft_size - 1)) != 0:
            raise ValueError(f"FFT size must be power of 2: {self.fft_size}")
        if self.hop_size <= 0:
            raise ValueError(f"Hop size must be positive: {self.hop_size}")
        if self.weight <= 0:
            raise ValueError(f"Weight must be positive: {self.weight}")


class OracleRing:
    """CircularBuffer (multi_resolution_fft.py:52-133): float32 ring; an oversize chunk keeps
    its last ``size`` samples; ``read_latest`` returns None until ``length`` samples exist."""

    def __init__(self, size: int):
        if size <= 0:
            raise ValueError("Buffer size must be positive")
        self.size = size
        self.buffer = np.zeros(size, dtype=np.float32)
        self.write_pos = 0
        self.samples_written = 0

    def write(self, data) -> bool:
        if data is None or len(data) == 0:
            return False
        n = len(data)
        if n >= self.size:                                   # :75-79
            self.buffer[:] = data[-self.size:]
            self.write_pos = 0
            self.samples_written = self.size
        else:
            if self.write_pos + n <= self.size:              # :82-83
                self.buffer[self.write_pos:self.write_pos + n] = data
            else:                                            # :84-88
                first = self.size - self.write_pos
                self.buffer[self.write_pos:] = data[:first]
                self.buffer[:n - first] = data[first:]
            self.write_pos = (self.write_pos + n) % self.size
            self.samples_written = min(self.samples_written + n, self.size)
        return True

    def read_latest(self, length: int) -> Optional[np.ndarray]:
        if length <= 0 or length > self.size:
            return None
        if self.samples_written < length:                    # :106-108
            return None
        out = np.zeros(length, dtype=np.float32)
        if self.write_pos >= length:
            out[:] = self.buffer[self.write_pos - length:self.write_pos]
        else:
            first = length - self.write_pos
            out[:first] = self.buffer[-first:]
            if self.write_pos > 0:
                out[first:] = self.buffer[:self.write_pos]
        return out


def psycho_weights(freqs: np.ndarray, freq_range, weight: float) -> np.ndarray:
    """_apply_psychoacoustic_weighting weight vector (multi_resolution_fft.py:304-326):
    ``config.weight`` everywhere, then inside ``freq_range`` multiplied by 1.8 (60-120 Hz),
    1.4 (200-400), 1.2 (2-5 kHz), 1.6 (20-80), overlapping masks multiply.  float32 in-place
    arithmetic as in the reference."""
    w = np.ones(len(freqs), dtype=np.float32)
    w.fill(weight)
    in_range = (freqs >= freq_range[0]) & (freqs <= freq_range[1])
    w[in_range & (freqs >= 60) & (freqs <= 120)] *= 1.8
    w[in_range & (freqs >= 200) & (freqs <= 400)] *= 1.4
    w[in_range & (freqs >= 2000) & (freqs <= 5000)] *= 1.2
    w[in_range & (freqs >= 20) & (freqs <= 80)] *= 1.6
    return w


class OracleMultiResFFT:
    """MultiResolutionFFT (multi_resolution_fft.py:135-408), per-chunk, one channel."""

    def __init__(self, sample_rate: int = 48000, max_freq: float = 20000, configs=None):
        if sample_rate <= 0:
            raise ValueError("Sample rate must be positive")
        if max_freq <= 0 or max_freq > sample_rate / 2:
            raise ValueError("Max frequency must be positive and <= Nyquist")
        self.sample_rate = sample_rate
        self.nyquist = sample_rate / 2
        self.max_freq = min(max_freq, self.nyquist)
        cfgs = DEFAULT_CONFIGS if configs is None else configs
        self.configs = [c if isinstance(c, OracleFFTConfig) else OracleFFTConfig(*c) for c in cfgs]
        self.windows = [multires_window(c.window_type, c.fft_size) for c in self.configs]
        # ring size: multi_resolution_fft.py:202
        self.rings = [OracleRing(max(c.fft_size * 2, c.fft_size + c.hop_size)) for c in self.configs]
        self.freq_arrays = [np.fft.rfftfreq(c.fft_size, 1 / sample_rate) for c in self.configs]

    def bin_weights(self, i: int) -> np.ndarray:
        c = self.configs[i]
        return psycho_weights(self.freq_arrays[i], c.freq_range, c.weight)

    def process_audio_chunk(self, chunk, apply_weighting: bool = True) -> Dict[int, np.ndarray]:
        """process_audio_chunk (:228-302) -> {config_index: magnitude float32[N/2+1]}."""
        if chunk is None or len(chunk) == 0:
            return {}
        out = {}
        for i, c in enumerate(self.configs):
            if not self.rings[i].write(chunk):
                continue
            audio = self.rings[i].read_latest(c.fft_size)
            if audio is None:
                continue
            windowed = audio * self.windows[i]                 # float32 * float32 (:268)
            mag = np.abs(np.fft.rfft(windowed))                # complex64 -> float32 under numpy>=2 (:272-273)
            if apply_weighting:
                mag = mag * self.bin_weights(i)[:len(mag)]     # (:329)
            out[i] = mag
        return out

    def combine(self, results: Dict[int, np.ndarray], target_bins: int = 1024):
        """combine_results_optimized (:335-408) -> (float32[T], float64[T])."""
        target_freqs = np.linspace(0, self.max_freq, target_bins)
        if not results:
            return np.zeros(target_bins), target_freqs        # (:347-349) float64 zeros
        combined = np.zeros(target_bins, dtype=np.float32)
        weight_sum = np.zeros(target_bins, dtype=np.float32)
        for i, magnitude in results.items():
            c = self.configs[i]
            freqs = self.freq_arrays[i]
            fr = c.freq_range
            valid = (freqs >= fr[0]) & (freqs <= fr[1])
            if not np.any(valid):
                continue
            vf, vm = freqs[valid], magnitude[valid]
            if len(vf) < 2:
                continue
            tmask = (target_freqs >= fr[0]) & (target_freqs <= fr[1])
            if not np.any(tmask):
                continue
            idx = np.where(tmask)[0]
            interp = np.interp(target_freqs[tmask], vf, vm)
            combined[idx] += interp * c.weight
            weight_sum[idx] += c.weight
        ok = weight_sum > 0
        combined[ok] /= weight_sum[ok]
        return combined.copy(), target_freqs


def combine_tables(sample_rate: int, max_freq: float, configs: Sequence[OracleFFTConfig], target_bins: int):
    """Index/fraction tables equivalent to the np.interp call in combine_results_optimized
    (multi_resolution_fft.py:366-391).  For resolution r and target bin t inside its closed
    freq_range: value = m[lo] + (m[lo+1]-m[lo])*frac with lo, frac as np.interp would pick
    them (xp = the valid FFT-bin frequencies).  Returns per-resolution (tidx int32[], lo int32[],
    frac float64[]).  Used by tests to cross-check the product's host-side table builder."""
    target_freqs = np.linspace(0, max_freq, target_bins)
    out = []
    for c in configs:
        freqs = np.fft.rfftfreq(c.fft_size, 1 / sample_rate)
        fr = c.freq_range
        valid = np.where((freqs >= fr[0]) & (freqs <= fr[1]))[0]
        tmask = (target_freqs >= fr[0]) & (target_freqs <= fr[1])
        tidx = np.where(tmask)[0].astype(np.int32)
        if len(valid) < 2 or len(tidx) == 0:
            out.append((np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float64)))
            continue
        vf = freqs[valid]
        x = target_freqs[tidx]
        j = np.searchsorted(vf, x, side="right") - 1          # vf[j] <= x
        lo = np.clip(j, 0, len(vf) - 2)
        frac = (x - vf[lo]) / (vf[lo + 1] - vf[lo])
        frac = np.where(j < 0, 0.0, frac)                      # left clamp -> fp[0]
        frac = np.where(j >= len(vf) - 1, 1.0, frac)           # right end / clamp -> fp[-1]
        out.append((tidx, (valid[0] + lo).astype(np.int32), frac))
    return out


# --------------------------------------------------------------------------------------
# batched FFT entry points (omega4/optimization/batched_fft_processor.py, gpu_accelerated_fft.py)
# --------------------------------------------------------------------------------------
def batched_fit(audio, fft_size: int) -> np.ndarray:
    """prepare_batch sizing (batched_fft_processor.py:136-139): keep the LAST fft_size samples
    or right-zero-pad."""
    audio = np.asarray(audio)
    if len(audio) > fft_size:
        return audio[-fft_size:]
    if len(audio) < fft_size:
        return np.pad(audio, (0, fft_size - len(audio)))
    return audio


def batched_fft_cpu(audio, fft_size: int, window_type: str = "hann"):
    """_process_size_group_cpu (batched_fft_processor.py:269-285): window*audio -> rfft -> abs;
    'frequencies' hard-coded to 48 kHz (:283).  dtype follows numpy promotion (float64 audio *
    float32 window -> float64 -> complex128)."""
    a = batched_fit(audio, fft_size)
    windowed = a * batched_window(window_type, fft_size)
    cplx = np.fft.rfft(windowed)
    return {"magnitude": np.abs(cplx), "complex": cplx,
            "frequencies": np.fft.rfftfreq(fft_size, 1 / 48000)}


def gpufft_compute_fft(audio, window_type: str = "hann"):
    """GPUAcceleratedFFT.compute_fft CPU branch (gpu_accelerated_fft.py:114-161), cache bypassed."""
    a = np.asarray(audio)
    n = len(a)
    windowed = a * gpufft_window(window_type, n)
    cplx = np.fft.rfft(windowed)
    return np.abs(cplx), cplx


def gpufft_multi_resolution(audio, resolutions: Dict[str, int], window_type: str = "hann"):
    """compute_multi_resolution_fft (gpu_accelerated_fft.py:179-254): last-N slice or right
    zero pad per size, then compute_fft; 'freqs' assumes 48 kHz (:210,246)."""
    a = np.asarray(audio)
    out = {}
    for name, n in resolutions.items():
        chunk = a[-n:] if len(a) >= n else np.pad(a, (0, n - len(a)))
        mag, cplx = gpufft_compute_fft(chunk, window_type)
        out[name] = {"magnitude": mag, "complex": cplx, "freqs": np.fft.rfftfreq(n, 1 / 48000)}
    return out


# --------------------------------------------------------------------------------------
# mel band mapping (omega4/optimization/freq_mapper.py)
# --------------------------------------------------------------------------------------
def mel_band_indices(sample_rate: int, fft_size: int, num_bars: int) -> List[Tuple[int, int]]:
    """PrecomputedFrequencyMapper._create_mel_band_mapping (freq_mapper.py:83-124)."""
    binw = sample_rate / fft_size
    hz_to_mel = lambda hz: 2595 * np.log10(1 + hz / 700)
    mel_to_hz = lambda mel: 700 * (10 ** (mel / 2595) - 1)
    mel_points = np.linspace(hz_to_mel(20), hz_to_mel(20000), num_bars + 1)
    freq_points = [mel_to_hz(m) for m in mel_points]
    freq_points[0] = max(20, freq_points[0])
    freq_points[-1] = min(20000, freq_points[-1])
    bands = []
    for i in range(num_bars):
        if i >= len(freq_points) - 1:
            break
        s = int(freq_points[i] / binw)
        e = int(freq_points[i + 1] / binw)
        if e <= s:
            e = s + 1
        s = max(0, min(s, fft_size // 2))
        e = max(s + 1, min(e, fft_size // 2 + 1))
        bands.append((s, e))
    return bands


def compensation_curve(sample_rate: int, fft_size: int) -> np.ndarray:
    """_compute_compensation_curve (freq_mapper.py:146-163)."""
    f = np.arange(fft_size // 2 + 1) * (sample_rate / fft_size)
    c = np.ones_like(f)
    for i, freq in enumerate(f):
        if freq > 0:
            if freq < 100:
                c[i] = 1.0 + (100 - freq) / 100 * 0.5
            elif freq < 1000:
                c[i] = 1.0
            elif freq < 4000:
                c[i] = 1.0 + (freq - 1000) / 3000 * 0.3
            else:
                c[i] = 1.3 - (freq - 4000) / 16000 * 0.5
    return c


def map_spectrum_to_bars(spectrum, bands, num_bars: int, comp: Optional[np.ndarray] = None) -> np.ndarray:
    """map_spectrum_to_bars (freq_mapper.py:165-196): optional compensation when lengths match,
    bar = mean(spectrum[s:e]); stops at the first band reaching past the spectrum."""
    spectrum = np.asarray(spectrum)
    out = np.zeros(num_bars, dtype=np.float32)
    if comp is not None and len(spectrum) == len(comp):
        spectrum = spectrum * comp
    for i, (s, e) in enumerate(bands):
        if i >= num_bars or e > len(spectrum):
            break
        out[i] = np.mean(spectrum[s:e]) if e > s else (spectrum[s] if s < len(spectrum) else 0)
    return out


def magnitude_to_db(x) -> np.ndarray:
    """Consumers' dB conversion, panels/spectrogram_waterfall.py:85: 20 log10(max(x, 1e-10))."""
    return 20.0 * np.log10(np.maximum(np.asarray(x, dtype=np.float64), 1e-10))


def magnitude_to_db_plus(x) -> np.ndarray:
    """The plugin spectrogram panel's form, plugins/panels/spectrogram.py:72: 20 log10(x + 1e-10)."""
    return 20.0 * np.log10(np.asarray(x, dtype=np.float64) + 1e-10)


def waterfall_freq_indices(sample_rate: int, fft_size: int, min_freq: float = 20, max_freq: float = 20000):
    """SpectrogramWaterfall._setup_frequency_mapping (panels/spectrogram_waterfall.py:55-69): first bin
    >= min_freq, first bin >= max_freq (the last bin when max_freq lies beyond Nyquist)."""
    bins = np.linspace(0, sample_rate / 2, fft_size // 2 + 1)
    lo = int(np.argmax(bins >= min_freq))
    hi = int(np.argmax(bins >= max_freq))
    if hi == 0:
        hi = len(bins) - 1
    return lo, hi


class OracleWaterfall:
    """SpectrogramWaterfall.update + _normalize_spectrum (panels/spectrogram_waterfall.py:71-121), data side:
    slice -> 20 log10(max(x, 1e-10)) -> (max, min) into a 100-entry history -> with auto gain the 95th
    percentile of the last 20 maxima / 5th percentile of the last 20 minima -> (dB + gain - floor) /
    (peak - floor) clipped to [0, 1] (zeros when the range is not positive).  float64 here; the
    reference computes the same in float32."""

    def __init__(self, sample_rate: int = 48000, fft_size: int = 2048):
        self.freq_indices = waterfall_freq_indices(sample_rate, fft_size)
        self.auto_gain = True
        self.gain_adjustment = 0.0
        self.peak_history = deque(maxlen=100)
        self.current_peak = 0.0
        self.current_floor = -80.0

    def update(self, fft_data) -> Optional[np.ndarray]:
        if fft_data is None or len(fft_data) == 0:
            return None
        lo, hi = self.freq_indices
        db = magnitude_to_db(np.asarray(fft_data)[lo:hi])
        self.peak_history.append((float(db.max()), float(db.min())))
        if self.auto_gain:
            recent = list(self.peak_history)[-20:]
            self.current_peak = float(np.percentile([p[0] for p in recent], 95))
            self.current_floor = float(np.percentile([p[1] for p in recent], 5))
        rng = self.current_peak - self.current_floor
        if rng > 0:
            return np.clip((db + self.gain_adjustment - self.current_floor) / rng, 0.0, 1.0)
        return np.zeros_like(db)


# --------------------------------------------------------------------------------------
# app-level spectrum post-processing (omega4_main.py:855-926, 992-1056) -- SURVEY.md section 8f rank 1
# --------------------------------------------------------------------------------------
def app_compensation_gains(freqs, content_type: str = "instrumental", vocal_suppression: float = 0.0) -> np.ndarray:
    """apply_frequency_compensation (omega4_main.py:855-926) applied to a vector of ones: the
    per-bin float32 gain the in-place ``compensated[mask] *= g`` statements leave behind."""
    f = np.asarray(freqs)
    c = np.ones(len(f), dtype=np.float32)
    if content_type == "vocal":                                   # :876-891
        c[f < 60] *= 0.15
        c[(f >= 60) & (f < 250)] *= 0.2
        c[(f >= 250) & (f < 500)] *= 0.6
        c[(f >= 500) & (f < 2000)] *= 1.5
    else:                                                         # :892-908
        c[f < 60] *= 0.8
        c[(f >= 60) & (f < 250)] *= 1.0
        c[(f >= 250) & (f < 500)] *= 1.1
        c[(f >= 500) & (f < 2000)] *= 0.85
    c[(f >= 2000) & (f < 6000)] *= 1.2                            # :911-912
    c[(f >= 6000) & (f < 10000)] *= 0.8                           # :915-916
    c[f >= 10000] *= 0.3                                          # :919-920
    if vocal_suppression > 0:                                     # :922-925
        c[(f >= 800) & (f < 4000)] *= (1.0 - vocal_suppression * 0.5)
    return c


def app_smoothing_factors(bands, sample_rate: int = SAMPLE_RATE, fft_size_base: int = FFT_SIZE_BASE) -> np.ndarray:
    """Per-bar smoothing factor of omega4_main.py:1043-1052 (band start bin -> Hz -> 0.6 / 0.75 / 0.85)."""
    out = np.empty(len(bands), dtype=np.float64)
    for i, (s, _e) in enumerate(bands):
        hz = s * sample_rate / fft_size_base
        out[i] = 0.6 if hz < 250 else (0.75 if hz < 2000 else 0.85)
    return out


class OracleSpectrumPost:
    """The block of ProfessionalLiveAudioAnalyzer.process_audio_spectrum between the combined
    spectrum and ``band_values`` (omega4_main.py:992-1056), one frame at a time:
    P98 normalisation x 0.8 -> frequency compensation -> optional max normalisation -> mel band
    mean -> sqrt -> clamp [0, 1] -> per-band exponential smoothing against the previous frame.
    ``self.freqs`` there is rfftfreq(FFT_SIZE_BASE) cut to the spectrum length (:168, :1006), and the
    band table is the one built for FFT_SIZE_BASE/2+1 bins (:171-178) applied to the (shorter)
    combined spectrum: the loop stops at the first band reaching past it (:1012-1013)."""

    def __init__(self, bars: int = BARS_DEFAULT, sample_rate: int = SAMPLE_RATE, fft_size_base: int = FFT_SIZE_BASE,
                 freq_compensation: bool = True, normalization: bool = False, smoothing: bool = True,
                 content_type: str = "instrumental", vocal_suppression: float = 0.0):
        self.bars = bars
        self.bands = mel_band_indices(sample_rate, fft_size_base, bars)
        self.freqs = np.fft.rfftfreq(fft_size_base, 1 / sample_rate)
        self.freq_compensation, self.normalization, self.smoothing = freq_compensation, normalization, smoothing
        self.content_type, self.vocal_suppression = content_type, vocal_suppression
        self.sf = app_smoothing_factors(self.bands, sample_rate, fft_size_base)
        self.prev = None

    def n_valid(self, spectrum_len: int) -> int:
        n = 0
        for _s, e in self.bands:
            if e > spectrum_len:
                break
            n += 1
        return min(n, self.bars)

    def process(self, spectrum) -> Tuple[np.ndarray, np.ndarray]:
        """-> (band_values after smoothing, peak_values = band values before smoothing)."""
        spectrum = np.asarray(spectrum)
        if np.max(spectrum) > 0:                                          # :992-998
            ref = np.percentile(spectrum, 98)
            if ref > 0:
                spectrum = spectrum / ref * 0.8
        if self.freq_compensation:                                        # :1001-1002
            gains = app_compensation_gains(self.freqs[:len(spectrum)], self.content_type, self.vocal_suppression)
            spectrum = spectrum * gains[:len(spectrum)]
        if self.normalization and np.max(spectrum) > 0:                   # :1005-1006
            spectrum = spectrum / np.max(spectrum)
        vals = []
        for s, e in self.bands:                                           # :1012-1032
            if e > len(spectrum):
                break
            v = np.mean(spectrum[s:e]) if e > s else (spectrum[s] if s < len(spectrum) else 0)
            if v > 0:
                v = max(0, min(1, np.sqrt(v)))
            vals.append(v)
        band = np.array(vals[:self.bars])
        peak = band.copy()
        if self.smoothing and self.prev is not None:                      # :1041-1054
            for i in range(len(band)):
                band[i] = self.prev[i] * self.sf[i] + band[i] * (1 - self.sf[i])
        self.prev = band.copy()
        return band, peak


# --------------------------------------------------------------------------------------
# bass zoom panel, data side (omega4/panels/bass_zoom.py) -- SURVEY.md section 8f rank 3
# --------------------------------------------------------------------------------------
def bass_mapping(sample_rate: int = 48000, fft_size: int = 8192):
    """setup_bass_mapping (bass_zoom.py:50-97) -> (freq_ranges, bin groups)."""
    freqs = np.fft.rfftfreq(fft_size, 1 / sample_rate)
    valid = [i for i, f in enumerate(freqs) if 20 <= f <= 200]
    ranges, groups = [], []
    per = max(1, len(valid) // 31)
    for i in range(0, len(valid), per):
        g = valid[i:min(i + per, len(valid))]
        if g:
            f0, f1 = freqs[g[0]], freqs[g[-1]]
            if ranges:
                f0 = max(f0, ranges[-1][1])
            ranges.append((f0, f1))
            groups.append(g)
    return ranges, groups


def bass_bars_step(audio, bar_values, ranges, groups, fft_size: int = 8192) -> np.ndarray:
    """_process_bass_detail_internal (bass_zoom.py:141-214) without the wall-clock peak hold:
    returns the new bar values (float32, like the reference's array)."""
    audio = np.asarray(audio)
    window = np.hanning(min(len(audio), fft_size))
    if len(audio) < fft_size:
        padded = np.zeros(fft_size)
        padded[:len(audio)] = audio * window
    else:
        padded = audio[:fft_size] * window
    mag = np.abs(np.fft.rfft(padded))
    bars = np.array(bar_values, dtype=np.float32).copy()
    raw, mx = {}, 0.0
    for i, g in enumerate(groups):
        if i >= len(bars):
            break
        if len(g) > 0:
            v = np.mean(mag[g])
            c = (ranges[i][0] + ranges[i][1]) / 2
            v *= 0.3 if c < 60 else 0.6 if c < 100 else 1.0 if c < 150 else 0.8
            raw[i] = v
            mx = max(mx, v)
    scale = (0.85 / mx) * (np.log10(max(1.0, mx * 10)) / 2.0) if mx > 0 else 1.0
    for i in raw:
        sv = raw[i] * scale
        cv = 0.7 + (sv - 0.7) * 0.3 if sv > 0.7 else sv
        bars[i] = bars[i] * 0.1 + cv * 0.9 if cv > bars[i] else bars[i] * 0.6 + cv * 0.4
        bars[i] = max(0.0, min(1.0, bars[i]))
    return bars


# --------------------------------------------------------------------------------------
# professional meters (omega4/panels/professional_meters.py)
# --------------------------------------------------------------------------------------
def butter2_highpass(fc: float, fs: float):
    """scipy.signal.butter(2, fc/(fs/2), 'high') == iirfilter(2, ..., btype='high', ftype='butter')
    (professional_meters.py:54,62-64).  Closed form of the bilinear transform with pre-warping:
    K = tan(pi fc/fs); b = [1,-2,1]/(1+sqrt2 K+K^2); a = [1, 2(K^2-1), 1-sqrt2 K+K^2]/(1+sqrt2 K+K^2)."""
    k = np.tan(np.pi * fc / fs)
    norm = 1.0 / (1.0 + np.sqrt(2.0) * k + k * k)
    b = np.array([1.0, -2.0, 1.0]) * norm
    a = np.array([1.0, 2.0 * (k * k - 1.0) * norm, (1.0 - np.sqrt(2.0) * k + k * k) * norm])
    return b, a


def k_weighting_coeffs(sample_rate: int):
    """create_k_weighting_filter (professional_meters.py:48-72): Butterworth-2 HPF @38 Hz and a
    second Butterworth-2 HPF @1500 Hz (the "shelf"); shelf_gain is computed but never used."""
    hp_b, hp_a = butter2_highpass(38.0, sample_rate)
    sh_b, sh_a = butter2_highpass(1500.0, sample_rate)
    return {"hp_b": hp_b, "hp_a": hp_a, "shelf_b": sh_b, "shelf_a": sh_a}


def lfilter_zi2(b, a) -> np.ndarray:
    """scipy.signal.lfilter_zi for a 2nd-order section: solve (I - companion(a)^T) zi = b[1:] - a[1:] b[0]."""
    b = np.asarray(b, dtype=np.float64) / a[0]
    a = np.asarray(a, dtype=np.float64) / a[0]
    m = np.array([[1.0 + a[1], -1.0], [a[2], 1.0]])
    rhs = np.array([b[1] - a[1] * b[0], b[2] - a[2] * b[0]])
    return np.linalg.solve(m, rhs)


def lfilter_tdf2(b, a, x: np.ndarray, zi: np.ndarray) -> np.ndarray:
    """scipy.signal.lfilter for order 2 (transposed direct form II), along the last axis, batched
    over leading axes: y = z1 + b0 x; z1 = z2 + b1 x - a1 y; z2 = b2 x - a2 y."""
    b0, b1, b2 = (np.asarray(b, dtype=np.float64) / a[0])
    a1, a2 = a[1] / a[0], a[2] / a[0]
    x = np.asarray(x, dtype=np.float64)
    y = np.empty_like(x)
    z1 = np.array(zi[..., 0], dtype=np.float64, copy=True)
    z2 = np.array(zi[..., 1], dtype=np.float64, copy=True)
    for n in range(x.shape[-1]):
        xn = x[..., n]
        yn = z1 + b0 * xn
        z1 = z2 + b1 * xn - a1 * yn
        z2 = b2 * xn - a2 * yn
        y[..., n] = yn
    return y


def odd_ext(x: np.ndarray, n: int) -> np.ndarray:
    """scipy.signal._arraytools.odd_ext along the last axis."""
    left = 2 * x[..., :1] - x[..., n:0:-1]
    right = 2 * x[..., -1:] - x[..., -2:-(n + 2):-1]
    return np.concatenate((left, x, right), axis=-1)


def filtfilt2(b, a, x: np.ndarray) -> np.ndarray:
    """scipy.signal.filtfilt(b, a, x) with defaults (padtype='odd', padlen=3*max(len(a),len(b))=9,
    method='pad'), restated per SURVEY.md section 8 a13; batched over leading axes."""
    padlen = 3 * max(len(a), len(b))
    zi = lfilter_zi2(b, a)
    ext = odd_ext(np.asarray(x, dtype=np.float64), padlen)
    y = lfilter_tdf2(b, a, ext, zi * ext[..., :1])
    y = lfilter_tdf2(b, a, y[..., ::-1], zi * y[..., -1:])
    return y[..., ::-1][..., padlen:-padlen]


def apply_k_weighting(frames: np.ndarray, coeffs) -> np.ndarray:
    """apply_k_weighting (professional_meters.py:129-153), batched over leading axes."""
    x = np.asarray(frames, dtype=np.float64)
    rms = np.sqrt(np.mean(x ** 2, axis=-1))
    f = filtfilt2(coeffs["hp_b"], coeffs["hp_a"], x)
    s = filtfilt2(coeffs["shelf_b"], coeffs["shelf_a"], f)
    out = f + (s - f) * 0.3
    out[rms < 1e-6] = 0.0                                    # (:132-134)
    return out


# --------------------------------------------------------------------------------------
# A / C / Z weighting (professional_meters.py:74-127, 155-229) -- SURVEY.md section 8f rank 2
# --------------------------------------------------------------------------------------
def butter_lowhigh(order: int, fc: float, fs: float, btype: str):
    """scipy.signal.butter(order, fc/(fs/2), btype) for order 1 or 2, closed form of the bilinear
    transform with pre-warping (K = tan(pi fc/fs))."""
    k = np.tan(np.pi * fc / fs)
    if order == 1:
        a = np.array([1.0, (k - 1.0) / (k + 1.0)])
        b = (np.array([1.0, -1.0]) if btype == "high" else np.array([k, k])) / (1.0 + k)
        return b, a
    norm = 1.0 / (1.0 + np.sqrt(2.0) * k + k * k)
    a = np.array([1.0, 2.0 * (k * k - 1.0) * norm, (1.0 - np.sqrt(2.0) * k + k * k) * norm])
    b = (np.array([1.0, -2.0, 1.0]) if btype == "high" else np.array([1.0, 2.0, 1.0]) * k * k) * norm
    return b, a


def a_weighting_sections(sample_rate: int):
    """create_a_weighting_filter (professional_meters.py:74-107): hp1, hp2, lp1, lp2 as (b, a) pairs."""
    nyq = sample_rate / 2
    f1, f2, f3, f4 = 20.598997, 107.65265, 737.86223, 12194.217
    return [butter_lowhigh(2, f1, sample_rate, "high"), butter_lowhigh(1, f2, sample_rate, "high"),
            butter_lowhigh(1, f3, sample_rate, "low"), butter_lowhigh(2, min(f4 / nyq, 0.99) * nyq, sample_rate, "low")]


def c_weighting_sections(sample_rate: int):
    """create_c_weighting_filter (professional_meters.py:109-127): hp, lp."""
    nyq = sample_rate / 2
    f1, f4 = 20.598997, 12194.217
    return [butter_lowhigh(2, f1, sample_rate, "high"), butter_lowhigh(2, min(f4 / nyq, 0.99) * nyq, sample_rate, "low")]


def filtfilt_any(b, a, x: np.ndarray) -> np.ndarray:
    """scipy.signal.filtfilt(b, a, x), defaults, for first- or second-order (b, a): padlen =
    3 * max(len(a), len(b)) (6 or 9); a first-order section is the biquad with b2 = a2 = 0."""
    padlen = 3 * max(len(a), len(b))
    b3 = np.concatenate([np.asarray(b, dtype=np.float64), np.zeros(3 - len(b))])
    a3 = np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(3 - len(a))])
    zi = lfilter_zi2(b3, a3)
    ext = odd_ext(np.asarray(x, dtype=np.float64), padlen)
    y = lfilter_tdf2(b3, a3, ext, zi * ext[..., :1])
    y = lfilter_tdf2(b3, a3, y[..., ::-1], zi * y[..., -1:])
    return y[..., ::-1][..., padlen:-padlen]


def apply_weighting(frames: np.ndarray, mode: str, sample_rate: int = 48000) -> np.ndarray:
    """apply_weighting (professional_meters.py:219-229) for 'K', 'A' (:155-192), 'C' (:194-217), 'Z'."""
    x = np.asarray(frames, dtype=np.float64)
    if mode == "K":
        return apply_k_weighting(x, k_weighting_coeffs(sample_rate))
    if mode == "Z" or mode not in ("A", "C"):
        return x
    rms = np.sqrt(np.mean(x ** 2, axis=-1))
    y = x.copy()
    for b, a in (a_weighting_sections(sample_rate) if mode == "A" else c_weighting_sections(sample_rate)):
        y = filtfilt_any(b, a, y)
    if mode == "A":
        y = y * 2.5                                              # (:190)
    y[rms < 1e-6] = 0.0                                          # (:158-160, :197-199)
    return y


def lufs_instantaneous_mode(frames: np.ndarray, mode: str, sample_rate: int = 48000) -> np.ndarray:
    """calculate_lufs core (:236-246) with the selected weighting."""
    w = apply_weighting(frames, mode, sample_rate)
    ms = np.mean(w ** 2, axis=-1)
    with np.errstate(divide="ignore"):
        l = -0.691 + 10.0 * np.log10(ms)
    return np.where(ms > 1e-10, l, -100.0)


def resample_fft(x: np.ndarray, factor: int = 4) -> np.ndarray:
    """scipy.signal.resample(x, factor*len(x)) for real x of even length (FFT method):
    X = rfft(x); X[N/2] *= 0.5; irfft(zero-padded X, factor*N) * factor."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    X = np.fft.rfft(x, axis=-1)
    Y = np.zeros(x.shape[:-1] + (factor * n // 2 + 1,), dtype=X.dtype)
    Y[..., : n // 2 + 1] = X
    if n % 2 == 0:
        Y[..., n // 2] *= 0.5
    return np.fft.irfft(Y, factor * n, axis=-1) * float(factor)


def true_peak_db(frames: np.ndarray, oversampling: int = 4) -> np.ndarray:
    """calculate_true_peak (professional_meters.py:283-299), batched: -100 if peak < 1e-10."""
    peak = np.max(np.abs(resample_fft(frames, oversampling)), axis=-1)
    with np.errstate(divide="ignore"):
        db = 20.0 * np.log10(peak)
    return np.where(peak < 1e-10, -100.0, db)


def lufs_instantaneous(frames: np.ndarray, coeffs) -> np.ndarray:
    """calculate_lufs core (professional_meters.py:237-246): K-weight -> mean square ->
    -0.691 + 10 log10 if > 1e-10 else -100."""
    w = apply_k_weighting(frames, coeffs)
    ms = np.mean(w ** 2, axis=-1)
    with np.errstate(divide="ignore"):
        l = -0.691 + 10.0 * np.log10(ms)
    return np.where(ms > 1e-10, l, -100.0)


class OracleMeterStats:
    """The deque statistics of calculate_lufs (professional_meters.py:20-25, 248-279)."""

    def __init__(self):
        self.mom = deque(maxlen=int(0.4 * 60))
        self.short = deque(maxlen=int(3.0 * 60))
        self.integ = deque(maxlen=int(60 * 60))
        self.peaks = deque(maxlen=int(1.0 * 60))
        self.gate = -70.0
        self.cur = {"momentary": -100.0, "short_term": -100.0, "integrated": -100.0,
                    "range": 0.0, "true_peak": -100.0}

    def push(self, lufs_inst: float, tp_db: float) -> Dict[str, float]:
        self.mom.append(lufs_inst)
        self.short.append(lufs_inst)
        self.integ.append(lufs_inst)
        self.cur["momentary"] = np.mean(self.mom)
        self.cur["short_term"] = np.mean(self.short)
        gated = [v for v in self.integ if v > self.gate]
        if gated:
            self.cur["integrated"] = np.mean(gated)
            self.cur["range"] = np.percentile(gated, 95) - np.percentile(gated, 10)
        else:
            self.cur["integrated"] = -100.0
            self.cur["range"] = 0.0
        self.peaks.append(tp_db)
        self.cur["true_peak"] = max(self.peaks)
        return self.cur


METER_KEYS = ("momentary", "short_term", "integrated", "range", "true_peak")


class OracleMetering:
    """ProfessionalMetering, K mode (professional_meters.py:13-299), per frame."""

    def __init__(self, sample_rate: int = 48000):
        self.sample_rate = sample_rate
        self.coeffs = k_weighting_coeffs(sample_rate)
        self.stats = OracleMeterStats()

    def calculate_lufs(self, frame) -> Dict[str, float]:
        if len(frame) == 0:
            return self.stats.cur
        frame = np.asarray(frame)
        l = float(lufs_instantaneous(frame[None, :], self.coeffs)[0])
        tp = float(true_peak_db(frame[None, :])[0])
        return self.stats.push(l, tp)


# --------------------------------------------------------------------------------------
# the frame schedule shared by oracle, CPU baseline and GPU
# --------------------------------------------------------------------------------------
def meter_frames(x: np.ndarray, hop: int = CHUNK_SIZE, window: int = FFT_SIZE_BASE):
    """Meter frames of one channel: hop k (0-based) ends at e=(k+1)*hop; frame = x[e-W:e] (float32)
    * np.hanning(W) (float64) as the app builds it (omega4_main.py:942-954), emitted once e >= W.
    Returns (first_hop, float64[n_frames, W])."""
    x = np.asarray(x, dtype=np.float32)
    n_hops = len(x) // hop
    first = (window + hop - 1) // hop - 1
    if n_hops <= first:
        return first, np.zeros((0, window), dtype=np.float64)
    idx = (np.arange(first, n_hops)[:, None] + 1) * hop - window + np.arange(window)[None, :]
    return first, x[idx].astype(np.float64) * np_hanning(window)[None, :]


def analyze_channel(x, sample_rate: int = 48000, configs=BASELINE_CONFIGS, hop: int = CHUNK_SIZE,
                    target_bins: int = BARS_DEFAULT, max_freq: float = MAX_FREQ,
                    meter_window: int = FFT_SIZE_BASE, apply_weighting: bool = True,
                    keep_magnitudes: bool = False, meter_batch: int = 512):
    """One channel through the whole path on the shared schedule.

    Per hop k: chunk x[k*hop:(k+1)*hop] -> OracleMultiResFFT.process_audio_chunk -> combine
    (zeros while no resolution has filled); meters on the last ``meter_window`` samples once
    available, otherwise the meters' initial values.  Returns dict with
    'combined' float32[K,T], 'meters' float64[K,5] (METER_KEYS order), 'lufs_inst' / 'tp_db'
    float64[K] (NaN before the first meter frame) and optionally 'magnitudes'
    {res: (first_hop, float32[n, N/2+1])}.
    """
    x = np.asarray(x, dtype=np.float32)
    n_hops = len(x) // hop
    mr = OracleMultiResFFT(sample_rate, max_freq, configs)
    combined = np.zeros((n_hops, target_bins), dtype=np.float32)
    mags = {i: [] for i in range(len(mr.configs))}
    for k in range(n_hops):
        res = mr.process_audio_chunk(x[k * hop:(k + 1) * hop], apply_weighting)
        if res:
            combined[k] = mr.combine(res, target_bins)[0]
        if keep_magnitudes:
            for i, m in res.items():
                mags[i].append(m)
    coeffs = k_weighting_coeffs(sample_rate)
    first, frames = meter_frames(x, hop, meter_window)
    lufs_inst = np.full(n_hops, np.nan)
    tp = np.full(n_hops, np.nan)
    for s in range(0, len(frames), meter_batch):
        blk = frames[s:s + meter_batch]
        lufs_inst[first + s:first + s + len(blk)] = lufs_instantaneous(blk, coeffs)
        tp[first + s:first + s + len(blk)] = true_peak_db(blk)
    stats = OracleMeterStats()
    meters = np.zeros((n_hops, 5), dtype=np.float64)
    for k in range(n_hops):
        if k >= first:
            stats.push(lufs_inst[k], tp[k])
        meters[k] = [stats.cur[key] for key in METER_KEYS]
    out = {"combined": combined, "meters": meters, "lufs_inst": lufs_inst, "tp_db": tp}
    if keep_magnitudes:
        out["magnitudes"] = {}
        for i, c in enumerate(mr.configs):
            fh = (c.fft_size + hop - 1) // hop - 1
            arr = np.stack(mags[i]) if mags[i] else np.zeros((0, c.fft_size // 2 + 1), np.float32)
            out["magnitudes"][i] = (fh, arr)
    return out
