"""GPU-backed drop-in for the data side of ``omega4.panels.professional_meters`` (reference file
omega4/panels/professional_meters.py): ``ProfessionalMetering`` (:13-299) and the non-drawing part
of ``ProfessionalMetersPanel`` (:302-397, 542-579).  pygame is not imported: drawing stays with the
reference's panel, which only reads the attributes kept here (lufs_info, level_history,
true_peak_history, loudness_range_history, peak_hold_value, peak_hold_counter, transient_info).

K-weighting (two zero-phase biquad passes as a block-parallel scan, fp64 state), the mean-square /
LUFS conversion, the exact 4x FFT-interpolated true peak and the deque statistics (M/S/I/LRA/TP
windows of 24/180/3600/60 frames) all run in libomega4_cuda.so; the deque contents live in a
device-format state vector carried between calls.  The A / C / Z weighting modes (:74-127, 155-229)
run through the same kernel as cascades of first/second-order zero-phase sections.  Frames must have
2048 samples (FFT_SIZE_BASE, what the app feeds); other lengths raise -- there is no CPU fallback.
"""
from __future__ import annotations

from collections import deque
from typing import Dict, Optional, Tuple

import numpy as np

from .. import _native as N
from ..plan import AnalysisPlan, BASELINE_CONFIGS, METER_KEYS

_PLANS: Dict[Tuple[int, int, str], AnalysisPlan] = {}


def _meter_plan(sample_rate: int, device: int, mode: str = "K") -> AnalysisPlan:
    """One plan per (sample rate, device, weighting mode): the weighting is plan state."""
    mode = mode if mode in ("K", "A", "C") else "Z"                 # apply_weighting's else branch (:228-229)
    key = (int(sample_rate), int(device), mode)
    p = _PLANS.get(key)
    if p is None:
        # the FFT part of the plan is irrelevant for the meters; any valid configuration will do
        cfg = [c for c in BASELINE_CONFIGS if c[0][1] <= sample_rate / 2]
        p = AnalysisPlan(sample_rate, cfg, 512, min(20000, sample_rate / 2), device=device)
        if mode != "K":
            p.set_weighting(mode)
        _PLANS[key] = p
    return p


class ProfessionalMetering:
    """Professional audio metering (LUFS, K-weighting, True Peak), GPU backed."""

    def __init__(self, sample_rate: int = 48000, device: int = 0):
        self.sample_rate = sample_rate
        self.device = device
        self._plan = _meter_plan(sample_rate, device)
        self._state = np.zeros((1, N.METER_STATE_DOUBLES), dtype=np.float64)
        self._fresh = True
        self._out = np.zeros(N.N_METERS, dtype=np.float32)
        self.weighting_mode = "K"
        self.gate_threshold = -70.0
        c = self._plan.kw_coeffs
        self.k_weighting_filter = {"hp_b": c[0:3].copy(), "hp_a": c[3:6].copy(), "shelf_b": c[6:9].copy(),
                                   "shelf_a": c[9:12].copy(), "shelf_gain": 10 ** (4.0 / 20)}
        self.current_lufs = {"momentary": -100.0, "short_term": -100.0, "integrated": -100.0,
                             "range": 0.0, "true_peak": -100.0}
        self.current_true_peak = -100.0

    # histories are views of the device-format state (oldest first), for code that inspects them
    @property
    def lufs_integrated_history(self):
        n = int(self._state[0, 0])
        return deque(self._state[0, 8:8 + n].tolist(), maxlen=3600)

    @property
    def lufs_short_term_history(self):
        return deque(list(self.lufs_integrated_history)[-180:], maxlen=180)

    @property
    def lufs_momentary_history(self):
        return deque(list(self.lufs_integrated_history)[-24:], maxlen=24)

    @property
    def peak_history(self):
        n = int(self._state[0, 1])
        return deque(self._state[0, 8 + 3600:8 + 3600 + n].tolist(), maxlen=60)

    def _mode_plan(self, mode: Optional[str] = None) -> AnalysisPlan:
        return _meter_plan(self.sample_rate, self.device, self.weighting_mode if mode is None else mode)

    def _weighted(self, audio_data: np.ndarray, mode: str) -> np.ndarray:
        _, _, w = self._mode_plan(mode).meter_frames_host(np.asarray(audio_data, dtype=np.float64), want_weighted=True)
        return w[0]

    def apply_k_weighting(self, audio_data: np.ndarray) -> np.ndarray:
        return self._weighted(audio_data, "K")

    def apply_a_weighting(self, audio_data: np.ndarray) -> np.ndarray:
        return self._weighted(audio_data, "A")

    def apply_c_weighting(self, audio_data: np.ndarray) -> np.ndarray:
        return self._weighted(audio_data, "C")

    def apply_weighting(self, audio_data: np.ndarray) -> np.ndarray:
        """(:219-229) 'K' / 'A' / 'C'; anything else is Z-weighting and returns the input itself."""
        if self.weighting_mode in ("K", "A", "C"):
            return self._weighted(audio_data, self.weighting_mode)
        return audio_data

    def calculate_true_peak(self, audio_data: np.ndarray, oversampling: int = 4) -> float:
        if len(audio_data) == 0:
            return -100.0
        if oversampling != 4:
            raise N.Omega4CudaError("only 4x oversampling is implemented on the GPU")
        f = np.asarray(audio_data, dtype=np.float64)[None, :]
        if f.shape[1] != self._plan.meter_window:
            raise N.Omega4CudaError(f"meter frames must have {self._plan.meter_window} samples")
        tp = np.empty(1, np.float64)
        rc = N.lib().omega4_meter_frames(self._plan.handle, None, N.MEM_HOST, N.ptr(np.ascontiguousarray(f)), 1,
                                         None, N.ptr(tp), None)
        N.check(rc, "omega4_meter_frames")
        return float(tp[0])

    def calculate_lufs(self, audio_data: np.ndarray) -> Dict[str, float]:
        """(:231-281) returns the same mutable dict object on every call, as the reference does."""
        if len(audio_data) == 0:
            return self.current_lufs
        f = np.ascontiguousarray(audio_data, dtype=np.float64)
        plan = self._mode_plan()
        if f.shape != (plan.meter_window,):
            raise N.Omega4CudaError(f"meter frames must have {plan.meter_window} samples (got {f.shape}); "
                                    "other lengths are not implemented on the GPU and there is no CPU fallback")
        plan.set_gate_threshold(self.gate_threshold)             # read on every call, as :267 does (plans are shared)
        # weighting + mean square, true peak and the deque statistics in one host round trip
        rc = N.lib().omega4_meter_update(plan.handle, f.ctypes.data, self._state.ctypes.data, 1 if self._fresh else 0,
                                         self._out.ctypes.data, None, None)
        N.check(rc, "omega4_meter_update")
        self._fresh = False
        for k, v in zip(METER_KEYS, self._out):
            self.current_lufs[k] = float(v)
        return self.current_lufs


class ProfessionalMetersPanel:
    """Data side of the professional meters panel (:302-397, 542-579)."""

    def __init__(self, sample_rate: int = 48000, device: int = 0):
        self.sample_rate = sample_rate
        self.metering = ProfessionalMetering(sample_rate, device)
        self.transient_info = {"attack_time": 0.0, "punch_factor": 0.0, "transients_detected": 0}
        self.level_history = deque(maxlen=600)
        self.histogram_bins = np.linspace(-60, 0, 61)
        self.peak_hold_time = 1.0
        self.peak_hold_samples = int(self.peak_hold_time * 60)
        self.peak_hold_value = -100.0
        self.peak_hold_counter = 0
        self.true_peak_history = deque(maxlen=300)
        self.use_gated_measurement = True
        self.loudness_range_history = deque(maxlen=300)

    def update(self, audio_data: np.ndarray):
        self.lufs_info = self.metering.calculate_lufs(audio_data)
        self.level_history.append(self.lufs_info["momentary"])
        self.loudness_range_history.append(self.lufs_info["range"])
        current_peak = self.lufs_info.get("true_peak", -100.0)
        self.true_peak_history.append(current_peak)
        if current_peak > self.peak_hold_value:
            self.peak_hold_value = current_peak
            self.peak_hold_counter = self.peak_hold_samples
        else:
            self.peak_hold_counter -= 1
            if self.peak_hold_counter <= 0:
                self.peak_hold_value = current_peak
        # transient statistics (:376-397) -- O(W) host bookkeeping on the caller's frame
        if len(audio_data) > 1:
            env = np.abs(audio_data)
            attacks = np.where(np.diff(env) > 0.1)[0]
            if len(attacks) > 0:
                self.transient_info["attack_time"] = (attacks[0] / self.sample_rate) * 1000
                rms = np.sqrt(np.mean(np.asarray(audio_data) ** 2))
                self.transient_info["punch_factor"] = float(np.max(env) / rms) if rms > 0 else 0.0
                self.transient_info["transients_detected"] = len(attacks)

    def get_results(self) -> Dict[str, object]:
        return {"lufs": self.lufs_info if hasattr(self, "lufs_info") else None, "transient": self.transient_info}

    def set_weighting(self, mode: str):
        if mode in ["K", "A", "C", "Z"]:
            self.metering.weighting_mode = mode

    def set_peak_hold_time(self, seconds: float):
        self.peak_hold_time = max(0.0, seconds)
        self.peak_hold_samples = int(self.peak_hold_time * 60)

    def toggle_gating(self):
        self.use_gated_measurement = not self.use_gated_measurement

    def reset_peak_hold(self):
        self.peak_hold_value = -100.0
        self.peak_hold_counter = 0

    def get_level_histogram(self) -> Tuple[np.ndarray, np.ndarray]:
        if not self.level_history:
            return self.histogram_bins[:-1], np.zeros(len(self.histogram_bins) - 1)
        hist, _ = np.histogram(list(self.level_history), bins=self.histogram_bins)
        if hist.sum() > 0:
            hist = hist.astype(float) / hist.sum()
        return self.histogram_bins[:-1], hist
