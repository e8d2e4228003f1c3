#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- round-2 golden vectors, frozen from the UNMODIFIED reference in the
build container (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_r2.py

* meters_96k.npz  -- BASELINE configs[4] shape for the meters: ``ProfessionalMetering(96000)``
  (professional_meters.py:16-72, 129-153, 231-299) on eight channels of 96 kHz audio delivered in the
  capture side's s16le wire format (capture.py:571-574, x = int16 / 32768): plain programme, digital
  silence, hot / clipped, DC offset, very quiet, a 30 Hz tone under the 38 Hz high-pass, isolated
  impulses, a 10 kHz tone.  One meter update per 512-sample hop on the Hann-windowed last 2048 samples.
* waterfall.npz   -- ``SpectrogramWaterfall.update`` (panels/spectrogram_waterfall.py:71-121): slice,
  dB, 20-entry P95 / P5 auto-gain, normalise, clip; auto-gain on / off, with a gain adjustment; and the
  plugin panel's dB form ``20 log10(x + 1e-10)`` (plugins/panels/spectrogram.py:72).
* multires_96k_stress.npz -- six resolutions up to 32768 at 96 kHz on an adversarial stream (a click
  entering the 32768-sample window, silence, a full-scale tone): per-hop combined rows of the reference.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OMEGA4_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

sys.dont_write_bytecode = True
pg = types.ModuleType("pygame")
pg.Surface = object
pg.Rect = object
pg.font = types.SimpleNamespace(Font=object)
sys.modules["pygame"] = pg
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "audio-analyzer-omega_b200"))

from omega4.audio.multi_resolution_fft import MultiResolutionFFT, FFTConfig, WindowType  # noqa: E402
from omega4.panels.professional_meters import ProfessionalMetering  # noqa: E402
from omega4.panels.spectrogram_waterfall import SpectrogramWaterfall  # noqa: E402

from omega4_b200.batch.synth import synth_channel  # noqa: E402  (seeded generator only)

HOP, W = 512, 2048
KEYS = ("momentary", "short_term", "integrated", "range", "true_peak")
CFG5 = [((20, 60), 32768, 1024, 1.5), ((60, 200), 16384, 1024, 1.5), ((200, 1000), 8192, 512, 1.2),
        ((1000, 5000), 4096, 256, 1.0), ((5000, 12000), 2048, 256, 1.2), ((12000, 20000), 1024, 256, 1.5)]


def channels_96k(n_hops):
    """Eight float64 channels before quantisation to int16."""
    sr, n = 96000, n_hops * HOP
    t = np.arange(n) / sr
    rng = np.random.default_rng(96000)
    ch = [synth_channel(10, c, n, sr).astype(np.float64) for c in range(8)]
    ch[1][20 * HOP:30 * HOP] = 0.0                                   # digital silence: rms gate, -100 paths
    ch[2] *= 1.9                                                     # hot: saturates the int16 range
    ch[3] = 0.3 + 0.2 * ch[3]                                        # DC offset under the high-pass
    ch[4] *= 2e-4                                                    # a few LSB
    ch[5] = 0.8 * np.sin(2 * np.pi * 30.0 * t) + 0.01 * rng.standard_normal(n)   # below the 38 Hz corner
    ch[6] = np.zeros(n)
    ch[6][rng.integers(0, n, size=40)] = rng.uniform(-0.95, 0.95, size=40)      # isolated impulses
    ch[7] = 0.7 * np.sin(2 * np.pi * 10000.0 * t + 0.4) + 0.05 * rng.standard_normal(n)
    return np.stack(ch)


def gen_meters_96k():
    n_hops = 72
    x16 = np.clip(np.round(channels_96k(n_hops) * 32768.0), -32768, 32767).astype(np.int16)
    x = x16.astype(np.float32) / 32768.0                             # capture.py:574
    first = W // HOP - 1
    rows = np.zeros((8, n_hops - first, 5))
    inst = np.zeros((8, n_hops - first))
    tps = np.zeros((8, n_hops - first))
    kw = {}
    for c in range(8):
        m = ProfessionalMetering(96000)
        for k in range(first, n_hops):
            e = (k + 1) * HOP
            frame = x[c, e - W:e] * np.hanning(W)                    # float32 * float64 (omega4_main.py:953-954)
            tps[c, k - first] = m.calculate_true_peak(frame)
            r = m.calculate_lufs(frame)
            rows[c, k - first] = [r[key] for key in KEYS]
            inst[c, k - first] = m.lufs_momentary_history[-1]
            if k == 40:
                kw[f"kweighted_c{c}_h40"] = m.apply_k_weighting(frame)
    f = ProfessionalMetering(96000).k_weighting_filter
    np.savez_compressed(os.path.join(OUT, "meters_96k.npz"), x16=x16, sample_rate=96000, hop=HOP, window=W,
                        first_hop=first, meters=rows, lufs_inst=inst, tp_db=tps,
                        hp_b=f["hp_b"], hp_a=f["hp_a"], shelf_b=f["shelf_b"], shelf_a=f["shelf_a"], **kw)


def gen_waterfall():
    sr, n = 48000, 2048
    n_rows = 48
    x = synth_channel(4, 0, (n_rows + 3) * HOP, sr)
    x[20 * HOP:26 * HOP] = 0.0                                       # silence: every bin at the 1e-10 floor
    x[30 * HOP:34 * HOP] *= 1e-3
    spectra = np.zeros((n_rows, n // 2 + 1), np.float32)
    for k in range(n_rows):
        e = (k + 4) * HOP
        spectra[k] = np.abs(np.fft.rfft(x[e - n:e] * np.hanning(n).astype(np.float32)))
    freqs = np.fft.rfftfreq(n, 1 / sr)
    out = {"spectra": spectra, "sample_rate": sr, "fft_size": n}
    for tag, auto, gain in (("auto", True, 0.0), ("fixed", False, 0.0), ("auto_gain3", True, 3.0)):
        wf = SpectrogramWaterfall(sr, n)
        wf.auto_gain = auto
        wf.gain_adjustment = gain
        peaks, floors = [], []
        for k in range(n_rows):
            wf.update(spectra[k], freqs)
            peaks.append(wf.current_peak)
            floors.append(wf.current_floor)
        out[f"rows_{tag}"] = np.stack(list(wf.waterfall_data))          # native dtype of the reference
        out[f"peak_{tag}"] = np.array(peaks, dtype=np.float64)
        out[f"floor_{tag}"] = np.array(floors, dtype=np.float64)
        out["freq_indices"] = np.array(wf.freq_indices, dtype=np.int64)
    # plugins/panels/spectrogram.py:72 -- the plugin panel's conversion (the class needs the plugin
    # framework's lifecycle; the line is the whole arithmetic)
    out["plugin_db"] = 20 * np.log10(spectra[:8] + 1e-10)
    # a 96 kHz / 4096 instance: max_freq below Nyquist picks another slice
    wf = SpectrogramWaterfall(96000, 4096)
    out["freq_indices_96k_4096"] = np.array(wf.freq_indices, dtype=np.int64)
    wf2 = SpectrogramWaterfall(22050, 1024)                          # max_freq beyond Nyquist -> last bin
    out["freq_indices_22k_1024"] = np.array(wf2.freq_indices, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "waterfall.npz"), **out)


def gen_multires_96k_stress():
    sr = 96000
    n_hops = 150
    n = n_hops * HOP
    t = np.arange(n) / sr
    rng = np.random.default_rng(5150)
    x = 1e-4 * rng.standard_normal(n)
    x[70 * HOP + 17] += 0.9                                          # a click that enters, crosses and leaves every window
    x[90 * HOP:100 * HOP] = 0.0                                      # digital silence
    x[100 * HOP:] += 0.95 * np.sin(2 * np.pi * 41.0 * t[100 * HOP:])   # full-scale tone inside the 32768 range
    x[120 * HOP:] += 0.3 * np.sin(2 * np.pi * 130.0 * t[120 * HOP:])   # and one inside the 16384 range
    x = x.astype(np.float32)
    mr = MultiResolutionFFT(sample_rate=sr)
    mr.configs = [FFTConfig(fr, nn, h, w, WindowType.BLACKMAN) for (fr, nn, h, w) in CFG5]
    mr._setup_windows(); mr._setup_buffers(); mr._setup_frequency_arrays(); mr._setup_working_arrays()
    combined = np.zeros((n_hops, 512), np.float32)
    mags = {}
    for k in range(n_hops):
        res = mr.process_audio_chunk(x[k * HOP:(k + 1) * HOP])
        if res:
            combined[k] = mr.combine_results_optimized(res, target_bins=512)[0]
        if k in (75, 130, 149):
            for i in (0, 1):
                if i in res:
                    mags[f"mag_h{k}_r{i}"] = res[i].magnitude[:64].copy()     # the bins the combine step reads lie below 64
    full = {f"combined_h{k}": combined[k] for k in (75, 130, 149)}
    np.savez_compressed(os.path.join(OUT, "multires_96k_stress.npz"), x=x, sample_rate=sr, hop=HOP,
                        combined_low=combined[60:, :32], combined_first=60, **full, **mags)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_meters_96k()
    gen_waterfall()
    gen_multires_96k_stress()
    for f in ("meters_96k.npz", "waterfall.npz", "multires_96k_stress.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))
