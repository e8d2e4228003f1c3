"""GPU parity tests added in round 2: BASELINE configs[4] (96 kHz, 8 channels, six resolutions up to 32768)
as a verified configuration, the waterfall consumer, the bars kernel away from
its default shape, the gate threshold, argument validation of the Python binding.

Same bars as tests/test_gpu_parity.py: spectrum 0.01 dB, LUFS 0.01 LU, true peak 0.05 dBTP."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import assert_spectrum_close, assert_spectrum_close_per_transform, transform_peaks
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_DB, TOL_LU, TOL_TP = 0.01, 0.01, 0.05
HOP, W = 512, 2048


def _names(plan):
    return [n for n, _ in plan.kernel_times()]


# ------------------------------------------------------------------ configs[4]: meters at 96 kHz
def test_meters_96k_eight_channels_golden(golden):
    """ProfessionalMetering(96000) on eight channels (professional_meters.py:48-72, 129-153, 231-299), fed in
    the s16le wire format as one interleaved 7.1 stream -- through omega4_analyze_s16, through omega4_analyze
    on the decoded planar rows, with the float32-state (default) and the float64-state K-weighting kernels."""
    from omega4_b200.plan import AnalysisPlan, CONFIG5_96K
    g = golden("meters_96k.npz")
    x16 = g["x16"]                                                   # [8, n] planar int16
    f = int(g["first_hop"])
    plan = AnalysisPlan(96000, CONFIG5_96K, 512)
    inter = np.ascontiguousarray(x16.T)[None, :, :]                  # [1 stream, n, 8 channels]
    out = plan.analyze_s16_host(inter, want_combined=False, want_series=True)
    worst = {}
    for tag, o in (("s16", out),
                   ("planar", plan.analyze_host(x16.astype(np.float32) / 32768.0, want_combined=False, want_series=True))):
        dl = np.abs(o["lufs_inst"][:, f:] - g["lufs_inst"]).max()
        dt = np.abs(o["tp_db"][:, f:] - g["tp_db"]).max()
        dm = np.abs(o["meters"][:, f:].astype(np.float64) - g["meters"])
        assert dl <= TOL_LU and dt <= TOL_TP, (tag, dl, dt)
        assert dm[..., :4].max() <= TOL_LU and dm[..., 4].max() <= TOL_TP, (tag, dm.max(axis=(0, 1)))
        assert np.all(o["meters"][:, :f] == [-100, -100, -100, 0, -100])
        worst[tag] = (dl, dt)
    assert np.array_equal(out["meters"], o["meters"])                # the decode is exact: bit-identical rows
    # in practice far inside the gate, also with alpha four times smaller than at 48 kHz
    # (true peak: half-precision delayed phases on the batch path, ~0.01 dBTP; float32 with OMEGA4_FLAG_EXACT_TRUE_PEAK)
    assert worst["s16"][0] < 1e-4 and worst["s16"][1] < 0.03, worst
    from omega4_b200 import _native as N
    ex = plan.analyze_host(x16.astype(np.float32) / 32768.0, want_combined=False, want_series=True, flags=N.FLAG_EXACT_TRUE_PEAK)
    assert np.abs(ex["tp_db"][:, f:] - g["tp_db"]).max() < 1e-3
    assert np.all(out["lufs_inst"][1, 23:30] == -100.0)              # digital silence -> rms gate
    assert np.all(out["meters"][4, :, 2] == -100.0)                  # a few LSB never pass the -70 gate
    os.environ["OMEGA4_KW_F64"] = "1"
    try:
        o64 = plan.analyze_s16_host(inter, want_combined=False, want_series=True)
    finally:
        os.environ.pop("OMEGA4_KW_F64", None)
    assert np.abs(o64["lufs_inst"][:, f:] - g["lufs_inst"]).max() < 1e-5
    assert np.abs(o64["lufs_inst"][:, f:] - out["lufs_inst"][:, f:]).max() < 1e-4
    plan.close()


def test_meters_96k_through_the_shim(golden):
    from omega4_b200.panels.professional_meters import ProfessionalMetering
    g = golden("meters_96k.npz")
    f = int(g["first_hop"])
    for ch in (0, 2, 5):
        x = g["x16"][ch].astype(np.float32) / 32768.0
        m = ProfessionalMetering(96000)
        np.testing.assert_allclose(m.k_weighting_filter["hp_a"], g["hp_a"], rtol=0, atol=3e-15)
        for k in range(f, 40):
            frame = x[(k + 1) * HOP - W:(k + 1) * HOP] * np.hanning(W)
            d = m.calculate_lufs(frame)
            want = g["meters"][ch, k - f]
            got = np.array([d[key] for key in O.METER_KEYS])
            assert np.abs(got[:4] - want[:4]).max() <= TOL_LU and abs(got[4] - want[4]) <= TOL_TP, (ch, k)
            assert abs(m.calculate_true_peak(frame) - g["tp_db"][ch, k - f]) <= TOL_TP
        fr40 = x[41 * HOP - W:41 * HOP] * np.hanning(W)
        wk = m.apply_k_weighting(fr40)
        ref = g[f"kweighted_c{ch}_h40"]
        assert np.abs(wk - ref).max() <= 1e-7 * max(1.0, np.abs(ref).max())


# ------------------------------------------------------------------ configs[4]: six resolutions, adversarial stream
def test_multires_96k_adversarial_stream_golden(golden):
    """A click that crosses the 32768-sample window, digital silence and full-scale tones inside the 32768 /
    16384 ranges: the default path (exact-windowing hop-block DFT on the tensor cores, 64 / 32 block positions
    per bin) and the full FFT against the unmodified reference, per frame."""
    from omega4_b200 import _native as N
    from omega4_b200.plan import AnalysisPlan, CONFIG5_96K
    g = golden("multires_96k_stress.npz")
    x = g["x"]
    f0 = int(g["combined_first"])
    plan = AnalysisPlan(96000, CONFIG5_96K, 512)
    # largest magnitude of the transform behind every target bin, per hop (the floor of the gate counts from it)
    ref = O.analyze_channel(x, 96000, O.CONFIG5_96K, keep_magnitudes=True)
    np.testing.assert_allclose(ref["combined"][f0:, :32], g["combined_low"], rtol=3e-6, atol=1e-9)
    peaks = transform_peaks(ref, O.CONFIG5_96K, len(x) // HOP)
    got = {}
    # The tensor-core path carries an error floor of a few 1e-7 of the transform's largest magnitude (the
    # tensor-core accumulator truncates; tests/tools/tone_leak_probe.py, profiles/r02_tone_leak.txt): next to a
    # full-scale tone it holds 0.01 dB down to about -65 dB, the FFT kernel well below -80 dB.
    for tag, fl, floor_db in (("default", 0, -60.0), ("fft", N.FLAG_NO_BLOCKDFT, -80.0)):
        out = plan.analyze_host(x[None, :], want_meters=False, flags=N.FLAG_TIME_KERNELS | fl)
        names = _names(plan)
        if tag == "default":
            assert "blockdft_tc_gemm" in names and "blockdft_asm_32768" in names and "blockdft_asm_16384" in names, names
            assert "blockdft_asm_8192" in names                       # 256 + 512 + 1280 = 2048 GEMM columns
            assert not any(n in names for n in ("multires_fft_32768", "multires_fft_16384", "multires_fft_8192"))
        else:
            assert "multires_fft_32768" in names and not any(n.startswith("blockdft") for n in names)
        comb = out["combined"][0]
        got[tag] = comb
        low = comb[f0:, :32]
        assert np.array_equal(low == 0, g["combined_low"] == 0), tag
        # target bins 1..5 come from the 32768 / 16384 transforms (tensor cores by default)
        assert_spectrum_close_per_transform(low[:, 1:6], g["combined_low"][:, 1:6], peaks[f0:, 1:6], TOL_DB, floor_db,
                                            label=f"{tag} sparse bins")
        # the other resolutions: 8192 rides the tensor-core GEMM too in the default mode (bins 6 .. 25), the rest are
        # full FFT kernels.  Next to two strong pure tones the float32 noise of ANY fp32 transform -- the reference's
        # pocketfft included -- is ~1e-7 of the largest magnitude, i.e. 0.009 dB at -80 dB: the FFT kernels are
        # held to -70 dB on this adversarial stream, the tensor-core path to its -60 dB
        for k in (75, 130, 149):
            assert_spectrum_close_per_transform(comb[k, 6:], g[f"combined_h{k}"][6:], peaks[k, 6:], TOL_DB,
                                                max(floor_db, -70.0), label=f"{tag} hop {k}")
        assert np.all(comb[:63, 1] == 0) and comb[63, 1] > 0           # 32768 / 512 - 1: first filled hop
    scale = got["fft"].max(axis=1, keepdims=True) + 1e-20
    assert (np.abs(got["fft"] - got["default"]) / scale).max() < 1e-4
    # magnitudes of the two largest transforms (API-faithful output) at the probe hops
    out = plan.analyze_host(x[None, :], want_meters=False, want_magnitudes=True)
    for k in (75, 130, 149):
        for r in (0, 1):
            ref = g[f"mag_h{k}_r{r}"]
            m = out["magnitudes"][r][0, k, :64]
            assert np.abs(m - ref).max() <= 2e-6 * out["magnitudes"][r][0, k].max() + 1e-9, (k, r)
    plan.close()


@pytest.mark.parametrize("seed,mode", [(2, "tc"), (7, "tc"), (4, "fft")])
def test_random_sections_96k_six_resolutions(seed, mode):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "random_stress.py"), str(seed), mode, "config5"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "random stress ok" in out.stdout


def test_exact_windowing_edge_shapes_at_64_blocks(golden):
    """Tiles with carried history and ragged lengths through the 64 / 32-position groups of the fused epilogue:
    frames that straddle row tiles and calls that start inside a window must equal the one-shot result."""
    import torch
    from omega4_b200 import _native as N
    from omega4_b200.batch.driver import StreamBatch
    from omega4_b200.plan import AnalysisPlan, CONFIG5_96K
    g = golden("multires_96k_stress.npz")
    x = np.stack([g["x"], g["x"][::-1].copy()])
    n_hops = x.shape[1] // HOP
    plan = AnalysisPlan(96000, CONFIG5_96K, 512)
    one = plan.analyze_host(x, want_meters=False)["combined"]
    full = plan.analyze_host(x, want_meters=False, flags=N.FLAG_NO_BLOCKDFT)["combined"]
    sb = StreamBatch(plan, 2, 67)
    parts, done = [], 0
    for n in (67, 1, 30, 52):
        sb.tile_view(n).copy_(torch.from_numpy(x[:, done * HOP:(done + n) * HOP]))
        comb = torch.empty((2, n, 512), device="cuda")
        sb.push(n, combined=comb)
        parts.append(comb.cpu().numpy())
        done += n
    assert done == n_hops
    tiled = np.concatenate(parts, axis=1)
    scale = full.max(axis=2, keepdims=True) + 1e-20
    assert np.array_equal(tiled == 0, one == 0)
    assert (np.abs(tiled - one) / scale).max() < 2e-6              # same kernels, different tile cuts: fp32 summation order only
    # against the full FFT: the tensor-core path's error floor is a few 1e-6 of the TRANSFORM's largest magnitude; rows
    # whose own maximum lies 90 dB below it (the reversed stream starts with the full-scale tone while the two longest
    # transforms have not filled yet) are therefore measured against the channel's largest value
    assert (np.abs(one - full) / full.max(axis=(1, 2), keepdims=True)).max() < 2e-5
    plan.close()


# ------------------------------------------------------------------ section 8f rank 3: waterfall
def test_waterfall_golden(golden):
    """SpectrogramWaterfall.update (spectrogram_waterfall.py:71-121) frame by frame and as one batch, auto gain
    on / off, with a gain adjustment; state carried between calls."""
    from omega4_b200.panels.spectrogram_waterfall import SpectrogramWaterfall, spectrogram_db
    g = golden("waterfall.npz")
    spectra = g["spectra"]
    freqs = np.fft.rfftfreq(2048, 1 / 48000)
    for tag, auto, gain in (("auto", True, 0.0), ("fixed", False, 0.0), ("auto_gain3", True, 3.0)):
        wf = SpectrogramWaterfall(48000, 2048)
        assert wf.freq_indices == tuple(g["freq_indices"])
        wf.auto_gain, wf.gain_adjustment = auto, gain
        for k in range(len(spectra)):
            row = wf.update(spectra[k], freqs)
            assert abs(wf.current_peak - g[f"peak_{tag}"][k]) < 1e-3 and abs(wf.current_floor - g[f"floor_{tag}"][k]) < 1e-3, (tag, k)
            assert np.abs(row - g[f"rows_{tag}"][k]).max() <= 2e-5, (tag, k)
        assert len(wf.waterfall_data) == len(spectra) and len(wf.peak_history) == len(spectra)
        # one batch == frame by frame; tiles with carried state == one batch
        wb = SpectrogramWaterfall(48000, 2048)
        wb.auto_gain, wb.gain_adjustment = auto, gain
        rows, dbv = wb.update_batch(spectra, want_db=True)
        assert np.array_equal(rows, np.stack(list(wf.waterfall_data)))
        assert np.abs(dbv - 20 * np.log10(np.maximum(spectra[:, 1:854].astype(np.float64), 1e-10))).max() < TOL_DB
        wt = SpectrogramWaterfall(48000, 2048)
        wt.auto_gain, wt.gain_adjustment = auto, gain
        tiles = np.concatenate([wt.update_batch(spectra[s:s + 7]) for s in range(0, len(spectra), 7)])
        assert np.array_equal(tiles, rows)
    assert SpectrogramWaterfall(96000, 4096).freq_indices == tuple(g["freq_indices_96k_4096"])
    assert SpectrogramWaterfall(22050, 1024).freq_indices == tuple(g["freq_indices_22k_1024"])
    wf = SpectrogramWaterfall()
    assert wf.update(np.zeros(0), freqs) is None and wf.update(None, freqs) is None and len(wf.waterfall_data) == 0
    # the plugin panel's dB form (plugins/panels/spectrogram.py:72)
    assert np.abs(spectrogram_db(spectra[:8]) - g["plugin_db"]).max() < 1e-3
    assert np.abs(spectrogram_db(spectra[3]) - g["plugin_db"][3]).max() < 1e-3


def test_waterfall_device_batch_and_band_map_db_forms():
    """Device-resident rows of many channels against the oracle; omega4_band_map's two dB forms."""
    import torch
    from omega4_b200 import _native as N
    from omega4_b200.plan import band_map_host
    rng = np.random.default_rng(3)
    n_ch, n_rows, ln = 5, 33, 513
    spec = (np.abs(rng.standard_normal((n_ch, n_rows, ln))) * 10 ** rng.uniform(-6, 1, (n_ch, n_rows, 1))).astype(np.float32)
    spec[2, 10:14] = 0.0
    lo, hi = O.waterfall_freq_indices(48000, 1024)
    d_spec = torch.from_numpy(spec).cuda()
    norm = torch.empty((n_ch, n_rows, hi - lo), device="cuda")
    stat = torch.empty((n_ch, n_rows, 4), device="cuda")
    state = torch.zeros((n_ch, N.WATERFALL_STATE), device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for a, b in ((0, 20), (20, 33)):                                   # two tiles, state carried on the device
        sl = d_spec[:, a:b].contiguous()
        no, so = torch.empty((n_ch, b - a, hi - lo), device="cuda"), torch.empty((n_ch, b - a, 4), device="cuda")
        rc = N.lib().omega4_waterfall(0, s, N.MEM_DEVICE, sl.data_ptr(), n_ch, b - a, ln, lo, hi, 0, 1, 0.0, state.data_ptr(),
                                      1 if a == 0 else 0, None, no.data_ptr(), so.data_ptr())
        N.check(rc, "omega4_waterfall")
        norm[:, a:b], stat[:, a:b] = no, so
    torch.cuda.synchronize()
    got, st = norm.cpu().numpy(), stat.cpu().numpy()
    for ch in range(n_ch):
        wf = O.OracleWaterfall(48000, 1024)
        for k in range(n_rows):
            want = wf.update(spec[ch, k])
            assert abs(st[ch, k, 2] - wf.current_peak) < 1e-3 and abs(st[ch, k, 3] - wf.current_floor) < 1e-3
            assert np.abs(got[ch, k] - want).max() <= 2e-5, (ch, k)
    bands = O.mel_band_indices(48000, 1024, 64)
    row = spec[0, :4]
    lin = band_map_host(row, bands)
    assert np.abs(band_map_host(row, bands, db=True) - O.magnitude_to_db(lin)).max() < 1e-3
    assert np.abs(band_map_host(row, bands, db=2) - O.magnitude_to_db_plus(lin)).max() < 1e-3
    with pytest.raises(N.Omega4CudaError):
        N.check(N.lib().omega4_waterfall(0, None, N.MEM_HOST, spec.ctypes.data, 1, 1, ln, 5, 5, 0, 1, 0.0, None, 1, None,
                                         spec.ctypes.data, None), "omega4_waterfall")


# ------------------------------------------------------------------ bars kernel away from T = 512 / P98
@pytest.mark.parametrize("T,q", [(40, 98.0), (50, 98.0), (64, 90.0), (100, 50.0), (300, 98.0), (300, 10.0), (512, 50.0), (1000, 90.0)])
def test_bars_percentile_reference_for_any_shape(T, q):
    """np.percentile(spectrum, q) normalisation for spectrum lengths and percentiles where a lane runs out of
    candidates (the -1 pad used to win the unsigned-bit selection): band = sqrt(spectrum / P_q * 0.8) clamp."""
    import ctypes as C
    from omega4_b200 import _native as N
    rng = np.random.default_rng(T * 7 + int(q))
    spec = np.abs(rng.standard_normal((3, 9, T))).astype(np.float32)
    spec[0, 0, :] = 0.0                                               # all-zero row: no scaling
    spec[1, 2, 18:32] = 50.0                                          # the maxima inside one lane group
    spec[2, 4, : T // 2] = 0.0                                        # half of the row exactly zero
    bands = np.ascontiguousarray(np.stack([np.arange(T), np.arange(T) + 1], axis=1).astype(np.int32))   # identity bands
    d = N.BarsDesc()
    d.spectrum_len, d.n_bars = T, T
    d.bands = bands.ctypes.data_as(C.POINTER(C.c_int))
    d.gain = None
    d.smooth = None
    d.percentile, d.scale, d.normalize_max = q, 0.8, 0
    h = N.lib().omega4_bars_create(C.byref(d), 0)
    assert h, N.last_error()
    out = np.empty((3, 9, T), np.float32)
    N.check(N.lib().omega4_bars_run(h, None, N.MEM_HOST, spec.ctypes.data, 3, 9, None, 1, out.ctypes.data, None), "omega4_bars_run")
    N.lib().omega4_bars_destroy(h)
    for c in range(3):
        for k in range(9):
            row = spec[c, k]
            x = row
            if row.max() > 0:
                ref = np.percentile(row, q)
                if ref > 0:
                    x = row / ref * 0.8
            want = np.clip(np.sqrt(x), 0, 1)
            assert np.abs(out[c, k] - want).max() <= 1e-5, (T, q, c, k)


def test_bars_state_copy_with_many_segments():
    """Carried smoothing state with more than one segment per channel: segment 0 reads a private copy of the
    state the last segment overwrites."""
    import torch
    from omega4_b200.app.spectrum_post import SpectrumPostProcessor
    rng = np.random.default_rng(12)
    spec = np.abs(rng.standard_normal((3, 1100, 512))).astype(np.float32)
    post = SpectrumPostProcessor(512)
    post._ensure()
    ref = post.process_host(spec)
    d = torch.from_numpy(spec).cuda()
    state = torch.zeros((3, 1 + post.n_valid), device="cuda")
    a = torch.empty((3, 550, post.n_valid), device="cuda")
    b = torch.empty((3, 550, post.n_valid), device="cuda")
    post.process_device(d[:, :550].contiguous(), a, state=state, fresh=True)
    post.process_device(d[:, 550:].contiguous(), b, state=state, fresh=False)
    torch.cuda.synchronize()
    got = torch.cat([a, b], dim=1).cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-6
    post.close()


# ------------------------------------------------------------------ gate threshold, binding validation
def test_gate_threshold_is_honoured(plan48):
    from omega4_b200 import _native as N
    rng = np.random.default_rng(4)
    li = rng.uniform(-90, -10, 300)
    tp = rng.uniform(-60, 0, 300)
    for gate in (-70.0, -40.0):
        plan48.set_gate_threshold(gate)
        got = plan48.meter_stats_host(li, tp, fresh=True)[0]
        st = O.OracleMeterStats()
        st.gate = gate
        want = []
        for a, b in zip(li, tp):
            d = st.push(a, b)                                          # one push per frame (the dict is reused)
            want.append([d[k] for k in O.METER_KEYS])
        want = np.array(want)
        assert np.abs(got - want).max() <= 1e-4, gate
    plan48.set_gate_threshold(-70.0)
    from omega4_b200.panels.professional_meters import ProfessionalMetering
    m = ProfessionalMetering(48000)
    m.gate_threshold = -5.0                                           # everything gated out
    t = np.arange(W) / 48000.0
    d = m.calculate_lufs(0.3 * np.sin(2 * np.pi * 1000 * t) * np.hanning(W))
    assert d["integrated"] == -100.0 and d["range"] == 0.0 and d["momentary"] > -40


def test_binding_rejects_bad_buffers(plan48):
    import torch
    from omega4_b200 import _native as N
    x = torch.zeros((2, 20 * HOP), device="cuda")
    comb = torch.empty((2, 20, 512), device="cuda")
    with pytest.raises(N.Omega4CudaError):                            # one entry per resolution
        plan48.analyze_device(x, 20, combined=comb, magnitudes=[None, None])
    with pytest.raises(N.Omega4CudaError):                            # wrong dtype
        plan48.analyze_device(x, 20, combined=comb.double())
    with pytest.raises(N.Omega4CudaError):                            # too small
        plan48.analyze_device(x, 20, combined=comb[:, :10].contiguous())
    with pytest.raises(N.Omega4CudaError):                            # non-contiguous
        plan48.analyze_device(x, 20, combined=torch.empty((2, 20, 1024), device="cuda")[:, :, ::2])
    with pytest.raises(N.Omega4CudaError):                            # float32 state
        plan48.analyze_device(x, 20, combined=comb, meters=torch.empty((2, 20, 5), device="cuda"),
                              meter_state=torch.zeros((2, N.METER_STATE_DOUBLES), device="cuda"))
    with pytest.raises(N.Omega4CudaError):
        plan48.analyze_device(x, 21, combined=comb)                   # rows shorter than n_hops * hop
    # a state vector with garbage counts must not index out of range (clamped in the kernel)
    st = np.full((1, N.METER_STATE_DOUBLES), 1e9)
    out = plan48.meter_stats_host(np.full(10, -20.0), np.full(10, -3.0), state=st, fresh=False)
    assert out.shape == (1, 10, 5)


@pytest.fixture(scope="module")
def plan48():
    from omega4_b200.plan import AnalysisPlan, BASELINE_CONFIGS
    p = AnalysisPlan(48000, BASELINE_CONFIGS, 512)
    yield p
    p.close()


# ------------------------------------------------------------------ omega4_analyze_io: fused band values, one descriptor
def test_analyze_io_fused_band_values_match_the_two_step_path(plan48):
    """int16 host input -> band_values + meters without the combined spectrum ever leaving the device: identical
    to omega4_analyze_s16 followed by omega4_bars_run (same kernels, same order), host and device mode, with the
    smoothing and meter states carried over time tiles."""
    import torch
    from omega4_b200 import _native as N
    from omega4_b200.app.spectrum_post import SpectrumPostProcessor
    from omega4_b200.batch.synth import synth_streams
    n_hops, n_streams = 90, 3
    x = synth_streams(n_streams, 2, n_hops * HOP)                          # [3, 2, n]
    x16 = np.ascontiguousarray(np.clip(np.round(x.transpose(0, 2, 1) * 32767.0), -32768, 32767).astype(np.int16))   # [3, n, 2]
    two = plan48.analyze_s16_host(x16, want_combined=True, want_meters=True)
    post = SpectrumPostProcessor(512)
    want_bars, want_peaks = post.process_host(two["combined"], want_peaks=True)
    nv, n_ch = post.n_valid, n_streams * 2
    band = np.empty((n_ch, n_hops, nv), np.float32)
    peak = np.empty_like(band)
    met = np.empty((n_ch, n_hops, 5), np.float32)
    plan48.analyze_io(N.MEM_HOST, n_ch, n_hops, x16.strides[0] // 2, frames_s16=x16, n_interleaved=2, meters=met,
                      bars=post, band_values=band, peak_values=peak)
    assert np.array_equal(band, want_bars) and np.array_equal(peak, want_peaks) and np.array_equal(met, two["meters"])
    # float input, combined requested as well
    xf = np.ascontiguousarray(x16.transpose(0, 2, 1).reshape(n_ch, -1).astype(np.float32) / 32768.0)
    comb = np.empty((n_ch, n_hops, 512), np.float32)
    band2 = np.empty_like(band)
    plan48.analyze_io(N.MEM_HOST, n_ch, n_hops, xf.shape[1], samples=xf, combined=comb, bars=post, band_values=band2)
    assert np.array_equal(comb, two["combined"]) and np.array_equal(band2, want_bars)
    # device mode in two time tiles with carried states == one shot
    sb_state = torch.zeros((n_ch, 1 + nv), device="cuda")
    m_state = torch.zeros((n_ch, N.METER_STATE_DOUBLES), dtype=torch.float64, device="cuda")
    d16 = torch.from_numpy(x16).cuda()
    outs, mets = [], []
    hist = 0
    for a, b in ((0, 50), (50, 90)):
        n = b - a
        bv = torch.empty((n_ch, n, nv), device="cuda")
        mv = torch.empty((n_ch, n, 5), device="cuda")
        base = d16.data_ptr() + a * HOP * 2 * 2
        plan48.analyze_io(N.MEM_DEVICE, n_ch, n, d16.stride(0), frames_s16=base, n_interleaved=2, hist=a * HOP, meters=mv,
                          meter_state=m_state, bars=post, band_values=bv, bars_state=sb_state,
                          flags=(N.FLAG_FRESH_METERS | N.FLAG_FRESH_BARS) if a == 0 else 0)
        outs.append(bv)
        mets.append(mv)
    torch.cuda.synchronize()
    tiled = torch.cat(outs, dim=1).cpu().numpy()
    assert np.abs(tiled - want_bars).max() <= 1e-6                     # tile cut changes the GEMM's row tiling: fp32 order
    assert np.array_equal(torch.cat(mets, dim=1).cpu().numpy(), two["meters"])
    with pytest.raises(N.Omega4CudaError):
        plan48.analyze_io(N.MEM_HOST, n_ch, n_hops, xf.shape[1], samples=xf, frames_s16=x16, meters=met)
    with pytest.raises(N.Omega4CudaError):
        plan48.analyze_io(N.MEM_HOST, n_ch, n_hops, xf.shape[1], samples=xf, bars=post, band_values=None)
    post.close()


def test_streaming_shims_one_round_trip(golden):
    """process_audio_chunk hands back the combination with the magnitudes; combine_results_optimized returns it
    only for those very magnitudes (otherwise it recomputes), and calculate_lufs is one call."""
    from omega4_b200.audio.multi_resolution_fft import MultiResolutionFFT
    from omega4_b200.plan import AnalysisPlan, DEFAULT_CONFIGS
    g = golden("multires_default.npz")
    x = g["x"]
    mr = MultiResolutionFFT(48000)
    plan = AnalysisPlan(48000, DEFAULT_CONFIGS, 512)
    for k in range(12):
        res = mr.process_audio_chunk(x[k * HOP:(k + 1) * HOP])
        if not res:
            continue
        c1, f1 = mr.combine_results_optimized(res, target_bins=512)
        mags = [res[i].magnitude if i in res else None for i in range(4)]
        want = plan.combine_host(mags, 1)[0]
        assert np.array_equal(c1, want), k
        launches = mr._plan(True, 512).launches
        c2, _ = mr.combine_results_optimized(res, target_bins=512)          # cached: no kernel launch
        assert np.array_equal(c2, want) and mr._plan(True, 512).launches == launches
        res[max(res)].magnitude[200] *= 2.0                                 # caller edits a magnitude inside its range: recomputed
        c3, _ = mr.combine_results_optimized(res, target_bins=512)
        mags = [res[i].magnitude if i in res else None for i in range(4)]
        assert np.array_equal(c3, plan.combine_host(mags, 1)[0]) and not np.array_equal(c3, want)
        c4, _ = mr.combine_results_optimized(res, target_bins=256)          # another length: recomputed
        assert c4.shape == (256,)
        mr._combine_bins = 512
    plan.close()


def test_streaming_shim_at_96k_six_resolutions(golden):
    """MultiResolutionFFT(96000) with the six config-5 resolutions fed hop by hop (omega4_stream_hop, including the
    32768-point kernel and the CUDA-graph replay once every ring has filled) against the reference's rows."""
    from omega4_b200.audio.multi_resolution_fft import MultiResolutionFFT, FFTConfig
    g = golden("multires_96k.npz")
    x = g["x"]
    mr = MultiResolutionFFT(96000)
    mr.configs = [FFTConfig(tuple(r), int(n), int(h), float(w)) for r, n, h, w in
                  zip(g["cfg_ranges"], g["cfg_sizes"], g["cfg_hops"], g["cfg_weights"])]
    mr._setup_windows(); mr._setup_buffers(); mr._setup_frequency_arrays(); mr._setup_working_arrays()
    n_hops = len(x) // HOP
    for k in range(n_hops):
        res = mr.process_audio_chunk(x[k * HOP:(k + 1) * HOP])
        assert sorted(res) == [i for i in range(6) if g["present"][k, i]], k
        if k >= 60:
            comb, freqs = mr.combine_results_optimized(res, target_bins=512)
            assert_spectrum_close(comb, g["combined_tail"][k - 60], TOL_DB, label=f"96k shim hop {k}")
        if k == 79:
            for i in range(6):
                assert_spectrum_close(res[i].magnitude, g[f"mag_h79_r{i}"], TOL_DB, label=f"96k shim r{i}")
                assert res[i].magnitude.dtype == np.float32 and len(res[i].frequencies) == len(res[i].magnitude)
    mr.cleanup()


def test_single_frame_statistics_equal_the_walking_kernel(plan48, golden):
    """One frame per call on carried state (stats_push1_kernel, what calculate_lufs runs per application frame) ==
    the whole series in one call (stats_kernel), bit for bit, through window saturation (3600) and gate straddling;
    and through omega4_meter_update on real frames."""
    from omega4_b200 import _native as N
    g = golden("meters_stats.npz")
    li, tp = g["lufs_inst"], g["tp_db"]
    one = plan48.meter_stats_host(li, tp, fresh=True)[0]
    np.testing.assert_allclose(one, g["meters"], rtol=0, atol=TOL_LU)
    state = np.zeros((1, N.METER_STATE_DOUBLES))
    rows = np.empty_like(one)
    for k in range(len(li)):
        rows[k] = plan48.meter_stats_host(li[k:k + 1], tp[k:k + 1], state=state, fresh=(k == 0))[0, 0]
    assert np.array_equal(rows, one)
    assert int(state[0, 0]) == 3600 and int(state[0, 1]) == 60
    # mixed: a tile, single frames, a tile
    state = np.zeros((1, N.METER_STATE_DOUBLES))
    parts = [plan48.meter_stats_host(li[:3590], tp[:3590], state=state, fresh=True)[0]]
    for k in range(3590, 3620):
        parts.append(plan48.meter_stats_host(li[k:k + 1], tp[k:k + 1], state=state)[0])
    parts.append(plan48.meter_stats_host(li[3620:], tp[3620:], state=state)[0])
    assert np.array_equal(np.concatenate(parts), one)
    # omega4_meter_update on explicit frames == omega4_meter_frames + omega4_meter_stats
    gm = golden("meters_stream.npz")
    x = gm["x"]
    frames = np.stack([x[(k + 1) * HOP - W:(k + 1) * HOP] * np.hanning(W) for k in range(3, 60)])
    l2, t2, _ = plan48.meter_frames_host(frames)
    want = plan48.meter_stats_host(l2, t2, fresh=True)[0]
    st = np.zeros(N.METER_STATE_DOUBLES)
    out = np.zeros(5, np.float32)
    for k, fr in enumerate(frames):
        fr = np.ascontiguousarray(fr)
        N.check(N.lib().omega4_meter_update(plan48.handle, fr.ctypes.data, st.ctypes.data, 1 if k == 0 else 0, out.ctypes.data, None, None),
                "omega4_meter_update")
        # the true-peak kernel packs two frames per transform: a frame measured alone and the same frame packed with
        # its neighbour differ in the last float32 digit; the loudness columns do not depend on the pairing
        assert np.array_equal(out[:4], want[k, :4]) and abs(out[4] - want[k, 4]) <= 1e-5, k


# ------------------------------------------------------------------ hop-block operand scales written by the K-weighting kernel
def test_row_scales_from_the_kweighting_kernel_equal_the_separate_pass(plan48):
    """The tensor-core GEMM's per-row operand scales come from the K-weighting kernel of the same call (it has every
    hop block in registers) and from the stand-alone kernel only for the blocks before the first meter frame or the
    carried history; OMEGA4_NO_SCALE_FOLD=1 keeps the separate pass.  Same scales -> bit-identical spectra, with and
    without history, with rows of very different levels, and without meters (nothing to fold into)."""
    rng = np.random.default_rng(11)
    n_ch, n_hops = 5, 70
    from omega4_b200 import _native as N
    for hist in (0, 8192 - HOP):
        x = rng.standard_normal((n_ch, hist + n_hops * HOP)).astype(np.float32)
        x *= np.array([1.0, 1e-6, 30.0, 0.0, 0.2], dtype=np.float32)[:, None]
        x[4, hist + 20 * HOP: hist + 24 * HOP] = 0.0                       # all-zero hop blocks inside a live row
        a = plan48.analyze_host(x, hist_samples=hist, flags=N.FLAG_TIME_KERNELS)
        t_fold = dict(plan48.kernel_times())
        assert "blockdft_tc_gemm" in t_fold
        os.environ["OMEGA4_NO_SCALE_FOLD"] = "1"
        try:
            b = plan48.analyze_host(x, hist_samples=hist, flags=N.FLAG_TIME_KERNELS)
            assert "blockdft_row_scale" in dict(plan48.kernel_times())
        finally:
            os.environ.pop("OMEGA4_NO_SCALE_FOLD", None)
        assert np.array_equal(a["combined"], b["combined"]) and np.array_equal(a["meters"], b["meters"])
        c = plan48.analyze_host(x, hist_samples=hist, want_meters=False)
        assert np.array_equal(a["combined"], c["combined"])
