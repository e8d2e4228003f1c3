"""Per-call latency of the reference-shaped streaming shims (one 512-sample hop at a time, as the pygame
application drives them): MultiResolutionFFT.process_audio_chunk + combine_results_optimized (the
reference's own benchmark_multi_fft driver), ProfessionalMetering.calculate_lufs, SpectrumPostProcessor."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "audio-analyzer-omega_b200"))
from omega4_b200.audio.multi_resolution_fft import benchmark_multi_fft
from omega4_b200.panels.professional_meters import ProfessionalMetering
from omega4_b200.app.spectrum_post import SpectrumPostProcessor

r = benchmark_multi_fft(48000, 512, 500)
print(f"MultiResolutionFFT.process_audio_chunk + combine_results_optimized: {r['avg_time_ms']:.3f} ms per hop "
      f"({r['iterations_per_second']:.0f} hops/s; real time needs 93.75 hops/s per channel)")
m = ProfessionalMetering(48000)
x = (np.random.default_rng(0).standard_normal(2048) * 0.1) * np.hanning(2048)
for _ in range(10):
    m.calculate_lufs(x)
t0 = time.perf_counter()
for _ in range(300):
    m.calculate_lufs(x)
print(f"ProfessionalMetering.calculate_lufs (K-weighting + true peak + deque statistics): {(time.perf_counter() - t0) / 300 * 1e3:.3f} ms per frame")
p = SpectrumPostProcessor(512)
s = np.abs(np.random.default_rng(1).standard_normal(512)).astype(np.float32)
for _ in range(10):
    p.process(s)
t0 = time.perf_counter()
for _ in range(300):
    p.process(s)
print(f"SpectrumPostProcessor.process: {(time.perf_counter() - t0) / 300 * 1e3:.3f} ms per frame")
