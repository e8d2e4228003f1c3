"""GPU-backed drop-in for ``omega4.optimization.batched_fft_processor`` (reference file
omega4/optimization/batched_fft_processor.py): ``BatchedFFTProcessor`` and
``get_batched_fft_processor()`` -- the app's batched-FFT entry point (omega4_main.py:181, 960-967).

``prepare_batch`` / ``process_batch`` / ``distribute_results`` / ``get_result_for_panel`` keep the
reference's request bookkeeping; ``process_batch`` groups requests by (size, window) and runs each
group as ONE ``omega4_rfft_batch`` launch (window fused into the load, |X| into the store, one
D2H of magnitude + complex).  Result dtype follows the reference's CuPy path: float32 magnitude,
complex64 spectrum (:237,244); 'frequencies' stays hard-coded to 48 kHz (:230,257,283).
"""
from __future__ import annotations

import threading
import time
from collections import defaultdict
from typing import Any, Dict, Optional

import numpy as np

from .. import _native as N
from .. import tables
from ..plan import rfft_batch_host
from .gpu_accelerated_fft import get_gpu_fft_processor


class FFTRequest:
    def __init__(self, request_id: str, audio_data: np.ndarray, fft_size: int, window_type: str = "hann"):
        self.request_id = request_id
        self.audio_data = audio_data
        self.fft_size = fft_size
        self.window_type = window_type
        self.result = None
        self.completed = False


class BatchedFFTProcessor:
    def __init__(self, gpu_memory_limit_mb: int = 256, device: int = 0):
        N.require_device()
        self.device = device
        self.gpu_available = True
        self.gpu_memory_limit = gpu_memory_limit_mb * 1024 * 1024
        self.pending_requests: Dict[str, FFTRequest] = {}
        self.request_lock = threading.Lock()
        self.common_fft_sizes = [512, 1024, 2048, 4096, 8192, 16384]
        self.window_cache: Dict[Any, np.ndarray] = {}
        self.batch_times = []
        self.last_batch_size = 0
        self._seq = 0
        self.gpu_fft = get_gpu_fft_processor()

    def _get_window(self, size: int, window_type: str) -> np.ndarray:
        key = (size, window_type)
        if key not in self.window_cache:
            self.window_cache[key] = tables.batched_window(window_type, size)
        return self.window_cache[key]

    def prepare_batch(self, panel_id: str, audio_data: np.ndarray, fft_size: int, window_type: str = "hann") -> str:
        """(:119-146) keep the LAST fft_size samples or right-zero-pad; returns the request id."""
        self._seq += 1
        request_id = f"{panel_id}_{fft_size}_{time.time()}_{self._seq}"
        audio_data = np.asarray(audio_data)
        if len(audio_data) > fft_size:
            audio_data = audio_data[-fft_size:]
        elif len(audio_data) < fft_size:
            audio_data = np.pad(audio_data, (0, fft_size - len(audio_data)))
        with self.request_lock:
            self.pending_requests[request_id] = FFTRequest(request_id, audio_data, fft_size, window_type)
        return request_id

    def process_batch(self) -> int:
        """(:148-195) one kernel launch per (fft_size, window_type) group."""
        with self.request_lock:
            if not self.pending_requests:
                return 0
            groups = defaultdict(list)
            for req in self.pending_requests.values():
                if not req.completed:
                    groups[(req.fft_size, req.window_type)].append(req)
        t0 = time.perf_counter()
        total = 0
        for (n, wt), reqs in groups.items():
            batch = np.stack([np.asarray(r.audio_data, dtype=np.float32) for r in reqs])
            mag, cx = rfft_batch_host(batch, self._get_window(n, wt), want_complex=True, device=self.device)
            freqs = np.fft.rfftfreq(n, 1 / 48000)
            for i, r in enumerate(reqs):
                r.result = {"magnitude": mag[i], "complex": cx[i], "frequencies": freqs}
                r.completed = True
            total += len(reqs)
        self.batch_times.append((time.perf_counter() - t0) * 1000)
        if len(self.batch_times) > 60:
            self.batch_times.pop(0)
        self.last_batch_size = total
        return total

    def distribute_results(self) -> Dict[str, Dict[str, np.ndarray]]:
        out = {}
        with self.request_lock:
            for rid in [k for k, r in self.pending_requests.items() if r.completed]:
                out[rid] = self.pending_requests.pop(rid).result
        return out

    def get_result_for_panel(self, request_id: str) -> Optional[Dict[str, np.ndarray]]:
        with self.request_lock:
            r = self.pending_requests.get(request_id)
            if r and r.completed:
                del self.pending_requests[request_id]
                return r.result
        return None

    def get_performance_stats(self) -> Dict[str, Any]:
        stats = {
            "gpu_enabled": self.gpu_available,
            "avg_batch_time": np.mean(self.batch_times) if self.batch_times else 0,
            "max_batch_time": max(self.batch_times) if self.batch_times else 0,
            "last_batch_size": self.last_batch_size,
            "pending_requests": len(self.pending_requests),
        }
        info = self.gpu_fft.get_gpu_memory_info()
        if info.get("available"):
            stats["gpu_memory_used_mb"] = info["used_mb"]
            stats["gpu_memory_total_mb"] = info["total_mb"]
        return stats

    def clear_gpu_cache(self):
        self.gpu_fft.clear_cache()

    def shutdown(self):
        self.clear_gpu_cache()
        self.window_cache.clear()
        self.pending_requests.clear()


_batched_fft_instance = None


def get_batched_fft_processor() -> BatchedFFTProcessor:
    global _batched_fft_instance
    if _batched_fft_instance is None:
        _batched_fft_instance = BatchedFFTProcessor()
    return _batched_fft_instance
