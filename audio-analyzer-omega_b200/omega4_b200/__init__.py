"""omega4_b200 -- B200-native (sm_100a) implementation of OMEGA-4's per-frame analysis hot path.

Layout mirrors the reference modules it stands in for:

    omega4_b200.audio.multi_resolution_fft      <- omega4/audio/multi_resolution_fft.py
    omega4_b200.optimization.batched_fft_processor / gpu_accelerated_fft / freq_mapper
    omega4_b200.panels.professional_meters      <- omega4/panels/professional_meters.py (data side)
    omega4_b200.batch                            headless multi-stream driver + 8-GPU partitioner
    omega4_b200.plan / tables / _native          plan object, host tables, ctypes binding

All arithmetic on the path runs in ``libomega4_cuda.so`` (csrc/, C ABI in include/omega4_cuda.h).
There is no CPU fallback: a missing library or device raises ``Omega4CudaError``.
"""
from ._native import Omega4CudaError, LIB_PATH  # noqa: F401

__all__ = ["Omega4CudaError", "LIB_PATH"]
