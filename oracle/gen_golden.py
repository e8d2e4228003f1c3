#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- freeze outputs of the UNMODIFIED reference into tests/golden/.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

It imports the reference's own classes (CPU path; cupy is absent) behind a 6-line pygame
stub, drives them with seeded inputs and stores input + output as small ``.npz`` fixtures.
The reference ships no golden vectors of its own, so these files ARE the
parity pin: ``tests/test_oracle_golden.py`` checks ``oracle/oracle_np.py`` against them and the
``-m gpu`` tests check the CUDA path against them.

Container that produced the committed fixtures: Python 3.12.3, numpy 2.3.5, scipy 1.18.1.
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
from omega4.optimization.batched_fft_processor import BatchedFFTProcessor  # noqa: E402
from omega4.optimization.gpu_accelerated_fft import GPUAcceleratedFFT  # noqa: E402
from omega4.optimization.freq_mapper import PrecomputedFrequencyMapper  # noqa: E402
from omega4.panels.professional_meters import ProfessionalMetering, ProfessionalMetersPanel  # noqa: E402

from omega4_b200.batch.synth import synth_channel  # noqa: E402  (seeded generator only)

HOP = 512
W = 2048
BASELINE = [((20, 200), 8192, 1024, 1.5), ((200, 1000), 4096, 512, 1.2),
            ((1000, 5000), 2048, 256, 1.0), ((5000, 20000), 1024, 256, 1.5)]


def make_multires(sr, configs, window_types=None):
    """MultiResolutionFFT with .configs overwritten and the four _setup_* re-run, exactly what
    SURVEY.md section 7 step 1 prescribes (the class itself is untouched)."""
    mr = MultiResolutionFFT(sample_rate=sr)
    if configs is not None:
        wts = window_types or [WindowType.BLACKMAN] * len(configs)
        mr.configs = [FFTConfig(fr, n, h, w, wt) for (fr, n, h, w), wt in zip(configs, wts)]
        mr._setup_windows()
        mr._setup_buffers()
        mr._setup_frequency_arrays()
        mr._setup_working_arrays()
    return mr


def run_multires(x, mr, target_bins, chunk=HOP, keep_hops=(), apply_weighting=True):
    n_hops = len(x) // chunk
    combined = np.zeros((n_hops, target_bins), dtype=np.float32)
    kept = {}
    present = np.zeros((n_hops, len(mr.configs)), dtype=np.uint8)
    for k in range(n_hops):
        res = mr.process_audio_chunk(x[k * chunk:(k + 1) * chunk], apply_weighting=apply_weighting)
        for i in res:
            present[k, i] = 1
        if res:
            combined[k] = mr.combine_results_optimized(res, target_bins=target_bins)[0]
        if k in keep_hops:
            for i, r in res.items():
                kept[f"mag_h{k}_r{i}"] = r.magnitude
    return combined, present, kept


def gen_multires():
    sr = 48000
    n = 96 * HOP
    x = synth_channel(0, 0, n, sr)                       # sweep 20 Hz..20 kHz inside the clip + pink
    keep = (0, 1, 3, 7, 15, 16, 50, 95)
    combined, present, kept = run_multires(x, make_multires(sr, BASELINE), 512, keep_hops=keep)
    unweighted, _, kept_u = run_multires(x, make_multires(sr, BASELINE), 512, keep_hops=(50,),
                                         apply_weighting=False)
    np.savez_compressed(os.path.join(OUT, "multires_baseline.npz"), x=x, sample_rate=sr, hop=HOP,
                        target_bins=512, combined=combined, present=present,
                        combined_unweighted_h50=unweighted[50],
                        **kept, **{k + "_unweighted": v for k, v in kept_u.items()})

    # reference default configuration (4096/2048/1024/1024), default target_bins=1024, second channel
    x2 = synth_channel(3, 1, 48 * HOP, sr)
    combined, present, kept = run_multires(x2, make_multires(sr, None), 1024, keep_hops=(6, 7, 47))
    np.savez_compressed(os.path.join(OUT, "multires_default.npz"), x=x2, sample_rate=sr, hop=HOP,
                        target_bins=1024, combined=combined, present=present, **kept)

    # app-style feed: the app hands 2048-sample Hann-windowed float64 frames as "chunks"
    # (omega4_main.py:953-954,980) -> oversize-chunk path of CircularBuffer for the 1024 rings
    mr = make_multires(sr, None)
    frames = []
    outs = {}
    for k in range(3, 12):
        fr = x2[(k + 1) * HOP - W:(k + 1) * HOP] * np.hanning(W)
        frames.append(fr)
        res = mr.process_audio_chunk(fr, apply_weighting=True)
        outs[f"present_{k}"] = np.array(sorted(res.keys()), dtype=np.int32)
        outs[f"combined_{k}"] = mr.combine_results_optimized(res, target_bins=512)[0]
        if k == 11:
            for i, r in res.items():
                outs[f"mag_{k}_r{i}"] = r.magnitude
    np.savez_compressed(os.path.join(OUT, "multires_appfeed.npz"), frames=np.stack(frames),
                        first_hop=3, **outs)

    # window quirks: HANN -> np.hann AttributeError -> rectangular; BLACKMAN_HARRIS -> blackman; HAMMING
    wts = [WindowType.HANN, WindowType.HAMMING, WindowType.BLACKMAN_HARRIS, WindowType.BLACKMAN]
    mrq = make_multires(sr, BASELINE, wts)
    combined, present, kept = run_multires(x[:24 * HOP], mrq, 512, keep_hops=(23,))
    np.savez_compressed(os.path.join(OUT, "multires_windows.npz"), n_samples=24 * HOP,
                        window_types=np.array([w.value for w in wts]), combined_h23=combined[23],
                        **{f"window_r{i}": mrq.windows[i] for i in range(4)}, **kept)

    # 96 kHz, six resolutions up to 32768 (BASELINE config 5 shape; ranges chosen per SURVEY section 7)
    sr5 = 96000
    cfg5 = [((20, 60), 32768, 1024, 1.5), ((60, 200), 16384, 1024, 1.5), ((200, 1000), 8192, 512, 1.2),
            ((1000, 5000), 4096, 256, 1.0), ((5000, 12000), 2048, 256, 1.2), ((12000, 20000), 1024, 256, 1.5)]
    x5 = synth_channel(5, 2, 80 * HOP, sr5)
    combined, present, kept = run_multires(x5, make_multires(sr5, cfg5), 512, keep_hops=(79,))
    np.savez_compressed(os.path.join(OUT, "multires_96k.npz"), x=x5, sample_rate=sr5, hop=HOP,
                        target_bins=512, combined_tail=combined[60:], present=present,
                        cfg_ranges=np.array([c[0] for c in cfg5], dtype=np.float64),
                        cfg_sizes=np.array([c[1] for c in cfg5]), cfg_hops=np.array([c[2] for c in cfg5]),
                        cfg_weights=np.array([c[3] for c in cfg5]), **kept)


def gen_meters():
    sr = 48000
    # (a) stream-scheduled frames
    n_hops = 120
    x = synth_channel(1, 0, n_hops * HOP, sr)
    x[40 * HOP:48 * HOP] = 0.0                            # digital silence -> rms gate / -100 paths
    x[60 * HOP:70 * HOP] *= 1e-4                          # very quiet
    x[80 * HOP:90 * HOP] *= 1.9                           # hot -> true peak above 0 dBTP
    m = ProfessionalMetering(sr)
    rows, inst, tps = [], [], []
    kw = {}
    for k in range(n_hops):
        e = (k + 1) * HOP
        if e < W:
            continue
        frame = x[e - W:e] * np.hanning(W)                # float32 * float64 (omega4_main.py:953-954)
        tps.append(m.calculate_true_peak(frame))
        r = m.calculate_lufs(frame)
        rows.append([r[key] for key in ("momentary", "short_term", "integrated", "range", "true_peak")])
        inst.append(m.lufs_momentary_history[-1])
        if k in (10, 44, 65, 85):
            kw[f"kweighted_h{k}"] = m.apply_k_weighting(frame)
    np.savez_compressed(os.path.join(OUT, "meters_stream.npz"), x=x, sample_rate=sr, hop=HOP, window=W,
                        first_hop=W // HOP - 1, meters=np.array(rows), lufs_inst=np.array(inst),
                        tp_db=np.array(tps), **kw)

    # (b) long run of the deque statistics: frames = gain[k] * base[k % 4]
    rng = np.random.default_rng(20241218)
    t = np.arange(W) / sr
    base = np.stack([
        0.5 * np.sin(2 * np.pi * 1000 * t),
        0.4 * np.sin(2 * np.pi * 100 * t) + 0.1 * np.sin(2 * np.pi * 7000 * t),
        rng.standard_normal(W) * 0.2,
        np.sign(np.sin(2 * np.pi * 440 * t)) * 0.9,
    ]) * np.hanning(W)[None, :]
    n = 4200
    gains_db = rng.uniform(-30, 0, size=n)
    gains_db[500:560] = -200.0                             # below every gate
    gains_db[900:1000] = rng.uniform(-75, -55, size=100)   # straddles the -70 gate
    gains_db[2000:2100] = -np.inf                          # exact zeros
    gains_db[3000:3700] = rng.uniform(-12, -8, size=700)
    gains = np.where(np.isinf(gains_db), 0.0, 10 ** (gains_db / 20))
    m = ProfessionalMetering(sr)
    rows, inst, tps = [], [], []
    for k in range(n):
        frame = gains[k] * base[k % 4]
        tps.append(m.calculate_true_peak(frame))
        r = m.calculate_lufs(frame)
        rows.append([r[key] for key in ("momentary", "short_term", "integrated", "range", "true_peak")])
        inst.append(m.lufs_momentary_history[-1])
    np.savez_compressed(os.path.join(OUT, "meters_stats.npz"), base=base, gains=gains, sample_rate=sr,
                        meters=np.array(rows), lufs_inst=np.array(inst), tp_db=np.array(tps))

    # (c) coefficients, known answers and the reference's own demo signals (test_enhanced_meters.py)
    m = ProfessionalMetering(sr)
    f = m.k_weighting_filter
    m96 = ProfessionalMetering(96000).k_weighting_filter
    sine = 0.5 * np.sin(2 * np.pi * 1000 * np.arange(W) / sr) * np.hanning(W)
    r = ProfessionalMetering(sr).calculate_lufs(sine)
    sq480 = np.ones(480) * 0.9
    sq480[::2] *= -1
    sq2048 = np.ones(W) * 0.9
    sq2048[::2] *= -1
    np.savez_compressed(
        os.path.join(OUT, "meters_known.npz"),
        hp_b=f["hp_b"], hp_a=f["hp_a"], shelf_b=f["shelf_b"], shelf_a=f["shelf_a"],
        hp_b96=m96["hp_b"], hp_a96=m96["hp_a"], shelf_b96=m96["shelf_b"], shelf_a96=m96["shelf_a"],
        sine_frame=sine, sine_momentary=r["momentary"], sine_true_peak=r["true_peak"],
        sq480=sq480, sq480_tp=ProfessionalMetering(sr).calculate_true_peak(sq480),
        sq480_lufs=ProfessionalMetering(sr).calculate_lufs(sq480)["momentary"],
        sq2048=sq2048, sq2048_tp=ProfessionalMetering(sr).calculate_true_peak(sq2048),
        sq2048_lufs=ProfessionalMetering(sr).calculate_lufs(sq2048)["momentary"],
        empty_tp=ProfessionalMetering(sr).calculate_true_peak(np.zeros(0)),
        zeros_tp=ProfessionalMetering(sr).calculate_true_peak(np.zeros(W)),
        zeros_lufs=ProfessionalMetering(sr).calculate_lufs(np.zeros(W))["momentary"],
    )

    # (d) panel extras: peak hold + transient stats (professional_meters.py:348-397) -- "next" row f2
    p = ProfessionalMetersPanel(sr)
    xs = synth_channel(2, 0, 60 * HOP, sr)
    xs[30 * HOP:31 * HOP] += 0.8 * np.sign(np.sin(np.arange(HOP)))   # clicks -> transient branch
    hold, cnt, att, punch, ntr = [], [], [], [], []
    for k in range(3, 60):
        frame = xs[(k + 1) * HOP - W:(k + 1) * HOP] * np.hanning(W)
        p.update(frame)
        hold.append(p.peak_hold_value)
        cnt.append(p.peak_hold_counter)
        att.append(p.transient_info["attack_time"])
        punch.append(p.transient_info["punch_factor"])
        ntr.append(p.transient_info["transients_detected"])
    np.savez_compressed(os.path.join(OUT, "meters_panel.npz"), x=xs, first_hop=3, peak_hold=np.array(hold),
                        peak_hold_counter=np.array(cnt), attack_time=np.array(att),
                        punch_factor=np.array(punch), transients=np.array(ntr),
                        level_history=np.array(p.level_history), histogram=p.get_level_histogram()[1])


def gen_batched():
    rng = np.random.default_rng(7)
    out = {}
    proc = BatchedFFTProcessor()
    sizes = [16384, 4096, 2048, 2048, 1024]               # test_batched_fft_performance.py:64-70
    ids = []
    for j, n in enumerate(sizes):
        a = rng.standard_normal(n).astype(np.float32)
        out[f"in_{j}"] = a
        ids.append(proc.prepare_batch(f"panel{j}", a, n))
    assert proc.process_batch() == len(sizes)
    res = proc.distribute_results()
    for j, rid in enumerate(ids):
        out[f"mag_{j}"] = res[rid]["magnitude"]
        out[f"cplx_{j}"] = res[rid]["complex"]
        out[f"freq_{j}"] = res[rid]["frequencies"]
    # window types, truncate (keep last N) and zero-pad, float64 input as the app passes it
    cases = [("hann", 3000, 2048), ("hamming", 1500, 2048), ("blackman", 2048, 2048),
             ("none", 700, 1024), ("hann", 512, 512)]
    for j, (wt, ln, n) in enumerate(cases):
        a = rng.standard_normal(ln)                        # float64
        rid = proc.prepare_batch("w", a, n, wt)
        proc.process_batch()
        r = proc.get_result_for_panel(rid)
        out[f"w_in_{j}"] = a
        out[f"w_mag_{j}"] = r["magnitude"]
        out[f"w_cplx_{j}"] = r["complex"]
    out["w_cases"] = np.array([f"{wt}:{ln}:{n}" for wt, ln, n in cases])
    # the app's double-Hann main spectrum request (omega4_main.py:953-960)
    xa = synth_channel(0, 1, 4096, 48000)[-2048:]
    fr = xa * np.hanning(2048)
    rid = proc.prepare_batch("main_spectrum", fr, 2048)
    proc.process_batch()
    r = proc.distribute_results()[rid]
    out["app_in"] = fr
    out["app_mag"] = r["magnitude"]
    out["app_cplx"] = r["complex"]

    g = GPUAcceleratedFFT()
    a = rng.standard_normal(4096).astype(np.float32)
    for wt in ("hann", "hamming", "blackman"):
        mag, cplx = g.compute_fft(a, wt, True)
        out[f"g_mag_{wt}"] = mag
        out[f"g_cplx_{wt}"] = cplx
    out["g_in"] = a
    a2 = rng.standard_normal(3000).astype(np.float32)     # non power of two, window built on the fly
    g.clear_cache()
    mag, cplx = g.compute_fft(a2, "hann", True)
    out["g_in_3000"] = a2
    out["g_mag_3000"] = mag
    a3 = rng.standard_normal(6000).astype(np.float32)
    mres = g.compute_multi_resolution_fft(a3, {"bass": 8192, "mid": 4096, "high": 1024}, "hann")
    out["gm_in"] = a3
    for name, d in mres.items():
        out[f"gm_mag_{name}"] = d["magnitude"]
        out[f"gm_cplx_{name}"] = d["complex"]
        out[f"gm_freqs_{name}"] = d["freqs"]
    np.savez_compressed(os.path.join(OUT, "batched_fft.npz"), **out)


def gen_freq_mapper():
    out = {}
    combos = [(48000, 2048, 512), (48000, 4096, 512), (48000, 8192, 512), (48000, 1024, 512),
              (48000, 4096, 1024), (96000, 32768, 512), (48000, 2048, 64), (44100, 2048, 256)]
    rng = np.random.default_rng(11)
    for sr, n, bars in combos:
        fm = PrecomputedFrequencyMapper(sr, n, bars)
        key = f"{sr}_{n}_{bars}"
        out["bands_" + key] = np.array(fm.mapping.band_indices, dtype=np.int32)
        spec = np.abs(rng.standard_normal(n // 2 + 1)).astype(np.float32)
        out["spec_" + key] = spec
        out["bars_comp_" + key] = fm.map_spectrum_to_bars(spec, apply_compensation=True)
        out["bars_raw_" + key] = fm.map_spectrum_to_bars(spec, apply_compensation=False)
        out["comp_" + key] = fm.mapping.compensation_curve
        # the app applies 1025-bin indices to the 512-bin combined spectrum (omega4_main.py:1011-1013)
        short = spec[:512]
        out["bars_short_" + key] = fm.map_spectrum_to_bars(short, apply_compensation=True)
    out["combos"] = np.array([f"{a}_{b}_{c}" for a, b, c in combos])
    np.savez_compressed(os.path.join(OUT, "freq_mapper.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_multires()
    gen_meters()
    gen_batched()
    gen_freq_mapper()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
