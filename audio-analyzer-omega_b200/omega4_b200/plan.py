"""AnalysisPlan -- owns one ``omega4_plan`` of libomega4_cuda.so (one per GPU / configuration).

The plan is the data form of the reference objects the hot path is built from:
``MultiResolutionFFT(sample_rate, max_freq).configs`` + ``combine_results_optimized(target_bins)``
(omega4/audio/multi_resolution_fft.py:138-160, 335) and ``ProfessionalMetering(sample_rate)``
(omega4/panels/professional_meters.py:16-46).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from . import tables

#: reference defaults, multi_resolution_fft.py:149-154
DEFAULT_CONFIGS = (
    ((20, 200), 4096, 1024, 1.5, "blackman"),
    ((200, 1000), 2048, 512, 1.2, "blackman"),
    ((1000, 5000), 1024, 256, 1.0, "blackman"),
    ((5000, 20000), 1024, 256, 1.5, "blackman"),
)
#: BASELINE.json sizes 8192/4096/2048/1024 substituted in order
BASELINE_CONFIGS = (
    ((20, 200), 8192, 1024, 1.5, "blackman"),
    ((200, 1000), 4096, 512, 1.2, "blackman"),
    ((1000, 5000), 2048, 256, 1.0, "blackman"),
    ((5000, 20000), 1024, 256, 1.5, "blackman"),
)
#: BASELINE config 5: 96 kHz, six resolutions up to 32768 (ranges chosen per SURVEY.md section 7)
CONFIG5_96K = (
    ((20, 60), 32768, 1024, 1.5, "blackman"),
    ((60, 200), 16384, 1024, 1.5, "blackman"),
    ((200, 1000), 8192, 512, 1.2, "blackman"),
    ((1000, 5000), 4096, 256, 1.0, "blackman"),
    ((5000, 12000), 2048, 256, 1.2, "blackman"),
    ((12000, 20000), 1024, 256, 1.5, "blackman"),
)

METER_KEYS = ("momentary", "short_term", "integrated", "range", "true_peak")


@dataclass(frozen=True)
class ResSpec:
    freq_range: Tuple[float, float]
    fft_size: int
    hop_size: int
    weight: float
    window_type: str = "blackman"


def _as_spec(c) -> ResSpec:
    if isinstance(c, ResSpec):
        return c
    if hasattr(c, "freq_range"):                           # FFTConfig-like
        wt = getattr(c, "window_type", "blackman")
        wt = getattr(wt, "value", wt)
        return ResSpec(tuple(c.freq_range), int(c.fft_size), int(c.hop_size), float(c.weight), str(wt))
    c = tuple(c)
    wt = c[4] if len(c) > 4 else "blackman"
    return ResSpec(tuple(c[0]), int(c[1]), int(c[2]), float(c[3]), str(getattr(wt, "value", wt)))


class AnalysisPlan:
    """Device tables + kernels for one (sample_rate, resolutions, target_bins) configuration."""

    def __init__(self, sample_rate: int = 48000, configs: Sequence = BASELINE_CONFIGS, target_bins: int = 512,
                 max_freq: float = 20000, hop: int = 512, apply_weighting: bool = True, device: int = 0,
                 windows: Optional[Sequence[np.ndarray]] = None):
        if sample_rate <= 0:
            raise ValueError("Sample rate must be positive")
        if max_freq <= 0 or max_freq > sample_rate / 2:
            raise ValueError("Max frequency must be positive and <= Nyquist")
        self.sample_rate = int(sample_rate)
        self.max_freq = min(max_freq, sample_rate / 2)
        self.hop = int(hop)
        self.target_bins = int(target_bins)
        self.apply_weighting = bool(apply_weighting)
        self.device = int(device)
        self.specs: List[ResSpec] = [_as_spec(c) for c in configs]
        if not (1 <= len(self.specs) <= N.MAX_RES):
            raise ValueError(f"1..{N.MAX_RES} resolutions supported")
        self.sizes = [s.fft_size for s in self.specs]
        self.freq_arrays = [np.fft.rfftfreq(s.fft_size, 1 / self.sample_rate) for s in self.specs]
        self.windows = [np.ascontiguousarray(w, dtype=np.float32) for w in windows] if windows is not None else \
            [tables.multires_window(s.window_type, s.fft_size) for s in self.specs]
        self.bin_weights = [tables.psycho_weights(f, s.freq_range, s.weight) for f, s in zip(self.freq_arrays, self.specs)]
        self.combine = tables.combine_tables(self.sample_rate, self.max_freq, self.sizes,
                                             [s.freq_range for s in self.specs], self.target_bins)
        self.target_freqs = np.linspace(0, self.max_freq, self.target_bins)
        self.meter_window = N.METER_WINDOW
        self.hann64 = np.hanning(self.meter_window).astype(np.float64)
        self.kw_coeffs = tables.k_weighting_coeffs(self.sample_rate)
        self._handle = None
        self.weighting_mode = "K"
        self._create()

    # ------------------------------------------------------------------ lifecycle
    def _create(self):
        lib = N.lib()
        N.require_device()
        keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a

        d = N.PlanDesc()
        d.sample_rate, d.hop, d.n_res = self.sample_rate, self.hop, len(self.specs)
        d.fft_sizes = arr(self.sizes, np.int32).ctypes.data_as(C.POINTER(C.c_int))
        d.windows = arr(np.concatenate(self.windows), np.float32).ctypes.data_as(C.POINTER(C.c_float))
        if self.apply_weighting:
            d.bin_weights = arr(np.concatenate(self.bin_weights), np.float32).ctypes.data_as(C.POINTER(C.c_float))
        else:
            d.bin_weights = None
        d.target_bins = self.target_bins
        d.tb_count = arr([len(t[0]) for t in self.combine], np.int32).ctypes.data_as(C.POINTER(C.c_int))
        cat = lambda k, dt: arr(np.concatenate([t[k] for t in self.combine]) if self.combine else [], dt)
        d.tb_idx = cat(0, np.int32).ctypes.data_as(C.POINTER(C.c_int))
        d.tb_lo = cat(1, np.int32).ctypes.data_as(C.POINTER(C.c_int))
        d.tb_frac = cat(2, np.float32).ctypes.data_as(C.POINTER(C.c_float))
        d.res_weight = arr([s.weight for s in self.specs], np.float32).ctypes.data_as(C.POINTER(C.c_float))
        d.meter_window = self.meter_window
        d.meter_hann = arr(self.hann64, np.float64).ctypes.data_as(C.POINTER(C.c_double))
        d.kw_coeffs = arr(self.kw_coeffs, np.float64).ctypes.data_as(C.POINTER(C.c_double))
        d.gate_threshold = -70.0
        h = lib.omega4_plan_create(C.byref(d), self.device)
        if not h:
            raise N.Omega4CudaError(f"omega4_plan_create failed: {N.last_error()}")
        self._handle = h

    def close(self):
        if self._handle:
            N.lib().omega4_plan_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._handle:
            raise N.Omega4CudaError("plan is closed")
        return self._handle

    def set_weighting(self, mode: str = "K"):
        """ProfessionalMetering.weighting_mode for every later meter call on this plan:
        'K' (default), 'A', 'C' or 'Z' (professional_meters.py:219-229)."""
        prog = tables.weighting_program(mode, self.sample_rate)
        w = N.Weighting()
        w.n_sections = len(prog["sections"])
        for i, (b, a) in enumerate(prog["sections"]):
            w.order[i] = len(a) - 1
            for j in range(len(b)):
                w.b[i][j] = float(b[j])
            for j in range(len(a)):
                w.a[i][j] = float(a[j])
        w.blend, w.rms_gate, w.gain = int(prog["blend"]), int(prog["rms_gate"]), float(prog["gain"])
        N.check(N.lib().omega4_plan_set_weighting(self.handle, C.byref(w)), "omega4_plan_set_weighting")
        self.weighting_mode = mode

    def set_gate_threshold(self, gate_threshold: float = -70.0):
        """ProfessionalMetering.gate_threshold (professional_meters.py:36, :267) for every later call."""
        N.check(N.lib().omega4_plan_set_gate_threshold(self.handle, float(gate_threshold)), "omega4_plan_set_gate_threshold")

    def _mag_ptrs(self, magnitudes, rows: int):
        """Pointer table of the optional per-resolution magnitude outputs: exactly one entry per resolution
        (None = not wanted), each float32, contiguous and large enough -- the C side indexes all n_res entries."""
        if magnitudes is None:
            return None
        if len(magnitudes) != len(self.sizes):
            raise N.Omega4CudaError(f"magnitudes needs {len(self.sizes)} entries (one per resolution), got {len(magnitudes)}")
        arr = (C.c_void_p * len(self.sizes))()
        for i, (m, n) in enumerate(zip(magnitudes, self.sizes)):
            arr[i] = N.checked(m, "float32", rows * (n // 2 + 1), f"magnitudes[{i}]")
        return arr

    # ------------------------------------------------------------------ helpers
    def first_hop(self, n: int, hist: int = 0) -> int:
        """First hop index at which a window of n samples is filled."""
        need = n - hist
        return 0 if need <= 0 else (need + self.hop - 1) // self.hop - 1

    @property
    def launches(self) -> int:
        return int(N.lib().omega4_plan_launches(self.handle))

    def kernel_times(self) -> List[Tuple[str, float]]:
        names = C.create_string_buffer(32 * 64)
        ms = (C.c_float * 64)()
        n = N.lib().omega4_plan_kernel_times(self.handle, names, ms, 64)
        out = []
        for i in range(max(n, 0)):
            out.append((names.raw[32 * i:32 * (i + 1)].split(b"\0", 1)[0].decode(), float(ms[i])))
        return out

    # ------------------------------------------------------------------ the hot path
    def analyze_host(self, samples: np.ndarray, hist_samples: int = 0, want_combined: bool = True,
                     want_magnitudes: bool = False, want_meters: bool = True, want_series: bool = False,
                     meter_state: Optional[np.ndarray] = None, flags: int = 0) -> Dict[str, object]:
        """Host buffers in, host buffers out (H2D/D2H inside the call).  ``samples`` float32
        [n_ch, hist_samples + n_hops*hop]; the first ``hist_samples`` columns are history."""
        x = np.ascontiguousarray(samples, dtype=np.float32)
        if x.ndim == 1:
            x = x[None, :]
        n_ch, total = x.shape
        n_hops = (total - hist_samples) // self.hop
        out: Dict[str, object] = {"n_hops": n_hops}
        if n_ch == 0 or n_hops <= 0:
            return out
        comb = np.empty((n_ch, n_hops, self.target_bins), np.float32) if want_combined else None
        mags = [np.empty((n_ch, n_hops, n // 2 + 1), np.float32) for n in self.sizes] if want_magnitudes else None
        met = np.empty((n_ch, n_hops, N.N_METERS), np.float32) if want_meters else None
        li = np.empty((n_ch, n_hops), np.float64) if want_series else None
        tp = np.empty((n_ch, n_hops), np.float64) if want_series else None
        base = x.ctypes.data + hist_samples * 4
        rc = N.lib().omega4_analyze(self.handle, None, N.MEM_HOST, base, x.strides[0] // 4, n_ch, n_hops,
                                    hist_samples, N.ptr(comb), self._mag_ptrs(mags, n_ch * n_hops), N.ptr(met),
                                    N.ptr(li), N.ptr(tp),
                                    N.checked(meter_state, "float64", n_ch * N.METER_STATE_DOUBLES, "meter_state"), flags)
        N.check(rc, "omega4_analyze")
        out.update(combined=comb, magnitudes=mags, meters=met, lufs_inst=li, tp_db=tp)
        return out

    def analyze_device(self, samples, n_hops: int, hist_samples: int = 0, combined=None, magnitudes=None,
                       meters=None, lufs_inst=None, tp_db=None, meter_state=None, flags: int = 0, stream=None):
        """Device tensors (torch) in / out, asynchronous on ``stream`` (default: torch's current).
        ``samples`` float32 [n_ch, >= hist_samples + n_hops*hop] with row stride samples.stride(0)."""
        import torch
        assert samples.is_cuda and samples.dtype == torch.float32 and samples.stride(-1) == 1
        n_ch = samples.shape[0]
        if stream is None:
            stream = torch.cuda.current_stream(samples.device).cuda_stream
        if samples.shape[1] < hist_samples + n_hops * self.hop:
            raise N.Omega4CudaError("samples rows are shorter than hist_samples + n_hops * hop")
        rows = n_ch * n_hops
        base = samples.data_ptr() + hist_samples * 4
        rc = N.lib().omega4_analyze(self.handle, stream, N.MEM_DEVICE, base, samples.stride(0), n_ch, n_hops,
                                    hist_samples, N.checked(combined, "float32", rows * self.target_bins, "combined"),
                                    self._mag_ptrs(magnitudes, rows),
                                    N.checked(meters, "float32", rows * N.N_METERS, "meters"),
                                    N.checked(lufs_inst, "float64", rows, "lufs_inst"), N.checked(tp_db, "float64", rows, "tp_db"),
                                    N.checked(meter_state, "float64", n_ch * N.METER_STATE_DOUBLES, "meter_state"), flags)
        N.check(rc, "omega4_analyze")

    def analyze_io(self, mem: int, n_ch: int, n_hops: int, stride: int, samples=None, frames_s16=None, n_interleaved: int = 1,
                   hist: int = 0, combined=None, meters=None, meter_state=None, bars=None, band_values=None, peak_values=None,
                   bars_state=None, flags: int = 0, stream=None):
        """``omega4_analyze_io``: the whole path with the application's post-processing fused behind the combine
        step.  ``samples`` (float32 planar rows) or ``frames_s16`` (interleaved int16) are numpy arrays / torch
        tensors / raw addresses of the first NEW sample; ``bars`` is a ``SpectrumPostProcessor`` (or its native
        handle).  Host mode is synchronous, device mode asynchronous on ``stream``."""
        io = N.IO()
        io.samples, io.frames_s16 = N.ptr(samples), N.ptr(frames_s16)
        io.n_interleaved, io.stride, io.n_ch, io.n_hops, io.hist = int(n_interleaved), int(stride), int(n_ch), int(n_hops), int(hist)
        rows = n_ch * n_hops
        io.combined = N.checked(combined, "float32", rows * self.target_bins, "combined")
        io.magnitudes = None
        io.meters = N.checked(meters, "float32", rows * N.N_METERS, "meters")
        io.lufs_inst = io.tp_db = None
        io.meter_state = N.checked(meter_state, "float64", n_ch * N.METER_STATE_DOUBLES, "meter_state")
        if bars is not None:
            handle = bars._ensure() if hasattr(bars, "_ensure") else bars
            nv = int(N.lib().omega4_bars_count(handle))
            io.bars = handle
            io.band_values = N.checked(band_values, "float32", rows * nv, "band_values")
            io.peak_values = N.checked(peak_values, "float32", rows * nv, "peak_values")
            io.bars_state = N.checked(bars_state, "float32", n_ch * (1 + nv), "bars_state")
        io.flags = int(flags)
        if mem == N.MEM_DEVICE and stream is None:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        N.check(N.lib().omega4_analyze_io(self.handle, stream, mem, C.byref(io)), "omega4_analyze_io")

    # ------------------------------------------------------------------ int16 wire format
    def analyze_s16_host(self, frames: np.ndarray, hist_frames: int = 0, want_combined: bool = True,
                         want_meters: bool = True, want_series: bool = False, flags: int = 0) -> Dict[str, object]:
        """Interleaved int16 frames as the capture side delivers them ("s16le", x = int16/32768,
        omega4/audio/capture.py:571-574): ``frames`` int16 [n_streams, hist_frames + n_hops*hop, C]
        (or [n_streams, n] for mono).  Outputs are indexed by planar channel stream*C + c."""
        f = np.ascontiguousarray(frames, dtype=np.int16)
        if f.ndim == 2:
            f = f[:, :, None]
        n_streams, total, il = f.shape
        n_hops = (total - hist_frames) // self.hop
        out: Dict[str, object] = {"n_hops": n_hops}
        if n_streams == 0 or n_hops <= 0:
            return out
        n_ch = n_streams * il
        comb = np.empty((n_ch, n_hops, self.target_bins), np.float32) if want_combined else None
        met = np.empty((n_ch, n_hops, N.N_METERS), np.float32) if want_meters else None
        li = np.empty((n_ch, n_hops), np.float64) if want_series else None
        tp = np.empty((n_ch, n_hops), np.float64) if want_series else None
        base = f.ctypes.data + hist_frames * il * 2
        rc = N.lib().omega4_analyze_s16(self.handle, None, N.MEM_HOST, base, f.strides[0] // 2, n_streams, il, n_hops,
                                        hist_frames, N.ptr(comb), None, N.ptr(met), N.ptr(li), N.ptr(tp), None, flags)
        N.check(rc, "omega4_analyze_s16")
        out.update(combined=comb, meters=met, lufs_inst=li, tp_db=tp)
        return out

    def analyze_s16_device(self, frames, n_hops: int, n_interleaved: int, hist_frames: int = 0, combined=None,
                           meters=None, meter_state=None, flags: int = 0, stream=None):
        """Device tensor int16 [n_streams, >= (hist_frames + n_hops*hop) * C] (interleaved)."""
        import torch
        assert frames.is_cuda and frames.dtype == torch.int16 and frames.stride(-1) == 1
        if stream is None:
            stream = torch.cuda.current_stream(frames.device).cuda_stream
        base = frames.data_ptr() + hist_frames * n_interleaved * 2
        rows = frames.shape[0] * n_interleaved * n_hops
        rc = N.lib().omega4_analyze_s16(self.handle, stream, N.MEM_DEVICE, base, frames.stride(0), frames.shape[0],
                                        n_interleaved, n_hops, hist_frames,
                                        N.checked(combined, "float32", rows * self.target_bins, "combined"), None,
                                        N.checked(meters, "float32", rows * N.N_METERS, "meters"), None, None,
                                        N.checked(meter_state, "float64", frames.shape[0] * n_interleaved * N.METER_STATE_DOUBLES,
                                                  "meter_state"), flags)
        N.check(rc, "omega4_analyze_s16")

    # ------------------------------------------------------------------ pieces used by the shims
    def combine_host(self, magnitudes: Sequence[Optional[np.ndarray]], n_rows: int = 1) -> np.ndarray:
        mags = [None if m is None else np.ascontiguousarray(m, dtype=np.float32).reshape(n_rows, -1) for m in magnitudes]
        for m, n in zip(mags, self.sizes):
            if m is not None and m.shape[1] != n // 2 + 1:
                raise ValueError("magnitude length does not match the resolution")
        out = np.empty((n_rows, self.target_bins), np.float32)
        rc = N.lib().omega4_combine(self.handle, None, N.MEM_HOST, self._mag_ptrs(mags, n_rows), n_rows, N.ptr(out))
        N.check(rc, "omega4_combine")
        return out

    def meter_frames_host(self, frames: np.ndarray, want_weighted: bool = False):
        f = np.ascontiguousarray(frames, dtype=np.float64)
        if f.ndim == 1:
            f = f[None, :]
        if f.shape[1] != self.meter_window:
            raise N.Omega4CudaError(f"meter frames must have {self.meter_window} samples (got {f.shape[1]}); "
                                    "other lengths are not implemented on the GPU and there is no CPU fallback")
        n = f.shape[0]
        li = np.empty(n, np.float64)
        tp = np.empty(n, np.float64)
        w = np.empty((n, self.meter_window), np.float64) if want_weighted else None
        rc = N.lib().omega4_meter_frames(self.handle, None, N.MEM_HOST, N.ptr(f), n, N.ptr(li), N.ptr(tp), N.ptr(w))
        N.check(rc, "omega4_meter_frames")
        return li, tp, w

    def meter_stats_host(self, lufs_inst: np.ndarray, tp_db: np.ndarray, state: Optional[np.ndarray] = None,
                         first_frame: int = 0, fresh: bool = False) -> np.ndarray:
        li = np.ascontiguousarray(lufs_inst, dtype=np.float64)
        tp = np.ascontiguousarray(tp_db, dtype=np.float64)
        if li.ndim == 1:
            li, tp = li[None, :], tp[None, :]
        n_ch, n = li.shape
        out = np.empty((n_ch, n, N.N_METERS), np.float32)
        rc = N.lib().omega4_meter_stats(self.handle, None, N.MEM_HOST, N.ptr(li), N.ptr(tp), n_ch, n, first_frame,
                                        N.checked(state, "float64", n_ch * N.METER_STATE_DOUBLES, "meter state"),
                                        N.ptr(out), 1 if fresh else 0)
        N.check(rc, "omega4_meter_stats")
        return out


# ---------------------------------------------------------------------- plan-less entry points
def rfft_batch_host(frames: np.ndarray, window: Optional[np.ndarray], want_complex: bool = True, device: int = 0):
    """Batched windowed rFFT of float32 rows -> (magnitude float32[B, n/2+1], complex64[B, n/2+1] | None)."""
    N.require_device()
    f = np.ascontiguousarray(frames, dtype=np.float32)
    if f.ndim == 1:
        f = f[None, :]
    b, n = f.shape
    mag = np.empty((b, n // 2 + 1), np.float32)
    cx = np.empty((b, n // 2 + 1), np.complex64) if want_complex else None
    w = None if window is None else np.ascontiguousarray(window, dtype=np.float32)
    rc = N.lib().omega4_rfft_batch(device, None, N.MEM_HOST, N.ptr(f), b, n, N.ptr(w), N.ptr(mag), N.ptr(cx))
    N.check(rc, "omega4_rfft_batch")
    return mag, cx


def band_map_host(spectrum: np.ndarray, bands: Sequence[Tuple[int, int]], comp: Optional[np.ndarray] = None,
                  db=False, device: int = 0) -> np.ndarray:
    """``db``: False = linear bars; True / 1 = 20 log10(max(x, 1e-10)) (panels/spectrogram_waterfall.py:85);
    2 = 20 log10(x + 1e-10) (plugins/panels/spectrogram.py:72)."""
    N.require_device()
    s = np.ascontiguousarray(spectrum, dtype=np.float32)
    if s.ndim == 1:
        s = s[None, :]
    rows, ln = s.shape
    b = np.ascontiguousarray(np.asarray(bands, dtype=np.int32).reshape(-1, 2))
    c = None if comp is None else np.ascontiguousarray(comp, dtype=np.float32)
    out = np.empty((rows, len(b)), np.float32)
    rc = N.lib().omega4_band_map(device, None, N.MEM_HOST, N.ptr(s), rows, ln, N.ptr(b), len(b), N.ptr(c), N.ptr(out),
                                 int(db))
    N.check(rc, "omega4_band_map")
    return out
