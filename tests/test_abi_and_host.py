"""CPU-only checks: the C-ABI library builds/loads and exports every symbol include/omega4_cuda.h
declares, it refuses to compute without a device (no CPU fallback), and the host-side tables the
product uploads agree with the oracle's restatement of the reference."""
import os
import re

import numpy as np
import pytest

from oracle import oracle_np as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as ge
    ge.build()
    from omega4_b200 import _native
    return _native


def test_library_exports_every_declared_symbol(native):
    hdr = open(os.path.join(ROOT, "include", "omega4_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(omega4_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    lib = native.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.omega4_abi_version() == native.ABI_VERSION == int(re.search(r"OMEGA4_ABI_VERSION (\d+)", hdr).group(1))
    assert int(re.search(r"OMEGA4_METER_STATE_DOUBLES \((.*?)\)", hdr).group(1).replace(" ", "").split("+")[1]) == 3600
    assert native.METER_STATE_DOUBLES == 8 + 3600 + 60


def test_only_sm100a_code_in_library(native):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_device(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from omega4_b200 import Omega4CudaError
    from omega4_b200.plan import AnalysisPlan, rfft_batch_host, band_map_host
    with pytest.raises(Omega4CudaError):
        AnalysisPlan()
    with pytest.raises(Omega4CudaError):
        rfft_batch_host(np.zeros((1, 1024), np.float32), None)
    with pytest.raises(Omega4CudaError):
        band_map_host(np.zeros(513, np.float32), [(0, 1)])
    from omega4_b200.audio.multi_resolution_fft import MultiResolutionFFT
    mr = MultiResolutionFFT()                       # construction is host-only bookkeeping
    assert mr.process_audio_chunk(np.zeros(512, np.float32)) == {}     # nothing filled yet: no GPU call
    with pytest.raises(Omega4CudaError):
        for _ in range(8):
            mr.process_audio_chunk(np.zeros(512, np.float32))
    from omega4_b200.panels.professional_meters import ProfessionalMetering
    with pytest.raises(Omega4CudaError):
        ProfessionalMetering()
    from omega4_b200.optimization.batched_fft_processor import BatchedFFTProcessor
    with pytest.raises(Omega4CudaError):
        BatchedFFTProcessor()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "audio-analyzer-omega_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)
                assert "/root/reference" not in src, os.path.join(dp, f)
                assert not re.search(r"^\s*(from|import)\s+scipy\b", src, flags=re.M), os.path.join(dp, f)


# ------------------------------------------------------------------ host tables vs oracle
def test_windows_weights_and_combine_tables_match_oracle():
    from omega4_b200 import tables
    for n in (512, 1024, 2048, 8192):
        for wt in ("blackman", "hann", "hamming", "blackman_harris", "other"):
            assert np.array_equal(tables.multires_window(wt, n), O.multires_window(wt, n))
        for wt in ("hann", "hamming", "blackman", "rect"):
            assert np.array_equal(tables.batched_window(wt, n), O.batched_window(wt, n))
            assert np.array_equal(tables.gpufft_window(wt, n), O.gpufft_window(wt, n))
    for sr, cfgs, T in ((48000, O.BASELINE_CONFIGS, 512), (48000, O.DEFAULT_CONFIGS, 1024), (44100, O.DEFAULT_CONFIGS, 300)):
        specs = [O.OracleFFTConfig(*c) for c in cfgs]
        ours = tables.combine_tables(sr, 20000, [c.fft_size for c in specs], [c.freq_range for c in specs], T)
        ref = O.combine_tables(sr, 20000, specs, T)
        for (ti, lo, fr), (rti, rlo, rfr) in zip(ours, ref):
            assert np.array_equal(ti, rti) and np.array_equal(lo, rlo)          # indices bit exact
            np.testing.assert_allclose(fr, rfr, atol=1e-7)
        for c in specs:
            f = np.fft.rfftfreq(c.fft_size, 1 / sr)
            assert np.array_equal(tables.psycho_weights(f, c.freq_range, c.weight), O.psycho_weights(f, c.freq_range, c.weight))


def test_mel_bands_and_meter_coefficients_match_oracle(golden):
    from omega4_b200 import tables
    g = golden("freq_mapper.npz")
    for combo in g["combos"]:
        sr, n, bars = (int(v) for v in str(combo).split("_"))
        assert np.array_equal(np.array(tables.mel_band_indices(sr, n, bars), np.int32), g["bands_" + str(combo)])
        f = np.arange(n // 2 + 1) * (sr / n)
        np.testing.assert_allclose(tables.compensation_curve(f), g["comp_" + str(combo)], atol=1e-15)
    gk = golden("meters_known.npz")
    c = tables.k_weighting_coeffs(48000)
    np.testing.assert_allclose(c, np.concatenate([gk["hp_b"], gk["hp_a"], gk["shelf_b"], gk["shelf_a"]]), atol=2e-15)
    c96 = tables.k_weighting_coeffs(96000)
    np.testing.assert_allclose(c96, np.concatenate([gk["hp_b96"], gk["hp_a96"], gk["shelf_b96"], gk["shelf_a96"]]), atol=2e-15)


def test_block_scan_formulation_of_filtfilt_matches_oracle():
    """numpy emulation of kweight_kernel's algorithm (32 lanes x 65 samples, direct-form-I
    zero-state sweep with the neighbour's input history, Kogge-Stone scan with Phi^(2^j),
    homogeneous correction, C^-14 virtual start for the backward pass) -- proves the scan tables
    build_biquad() derives are the right ones."""
    from omega4_b200 import tables
    L, NL, PAD, Wn = 65, 32, 9, 2048
    coef = tables.k_weighting_coeffs(48000)

    def biquad(b, a):
        a1, a2 = a[1], a[2]
        C = np.array([[-a1, -a2], [1.0, 0.0]])
        g = np.zeros((L, 2)); P = C.copy()
        for i in range(L):
            g[i] = P[0]
            if i + 1 < L:
                P = C @ P
        zi = O.lfilter_zi2(b, a)
        yinit = np.array([-zi[1] / a2, (-zi[0] + a1 * zi[1] / a2) / a2])
        return dict(b=b, a=a, g=g, phi=[np.linalg.matrix_power(P, 2 ** j) for j in range(5)],
                    cinv=np.linalg.inv(np.linalg.matrix_power(C, 14)), yinit=yinit)

    def one_pass(r, q, backward):
        b0, b1, b2 = q["b"]; a1, a2 = q["a"][1], q["a"][2]
        r = r.copy()
        x0 = r[31, 50] if backward else r[0, 0]
        if backward:
            r[31, 51:] = 0.0
        s_init = q["yinit"] * x0
        if backward:
            s_init = q["cinv"] @ s_init
        xo = r.copy()
        e = np.zeros((NL, 2))
        order = list(range(L - 1, -1, -1)) if backward else list(range(L))
        for l in range(NL):
            prev = l + 1 if backward else l - 1
            if 0 <= prev < NL:
                xm1, xm2 = (xo[prev, 0], xo[prev, 1]) if backward else (xo[prev, 64], xo[prev, 63])
            else:
                xm1 = xm2 = 0.0
            y1 = y2 = 0.0
            for i in order:
                x = r[l, i]
                y = -a1 * y1 + (-a2 * y2 + (b0 * x + (b1 * xm1 + b2 * xm2)))
                xm2, xm1, y2, y1 = xm1, x, y1, y
                r[l, i] = y
            e[l] = (y1, y2)
        pos = 31 - np.arange(NL) if backward else np.arange(NL)
        v = e.copy()
        v[np.where(pos == 0)[0][0]] += q["phi"][0] @ s_init
        for j in range(5):
            d = 1 << j
            nv = v.copy()
            for l in range(NL):
                if pos[l] >= d:
                    nv[l] = v[l] + q["phi"][j] @ v[l + d if backward else l - d]
            v = nv
        for l in range(NL):
            sin = s_init if pos[l] == 0 else v[l + 1 if backward else l - 1]
            for n, i in enumerate(order):
                r[l, i] += q["g"][n, 0] * sin[0] + q["g"][n, 1] * sin[1]
        return r

    def odd_pad(r):
        for j in range(PAD):
            r[0, j] = 2 * r[0, PAD] - r[0, 2 * PAD - j]
            r[31, 42 + j] = 2 * r[31, 41] - r[31, 40 - j]

    rng = np.random.default_rng(9)
    t = np.arange(Wn) / 48000
    xh = (0.5 * np.sin(2 * np.pi * 23 * t + 0.3) + 0.1 * rng.standard_normal(Wn)) * np.hanning(Wn)
    r = np.zeros((NL, L)); r.reshape(-1)[PAD:PAD + Wn] = xh
    hp = biquad(coef[0:3], coef[3:6]); sh = biquad(coef[6:9], coef[9:12])
    odd_pad(r)
    r = one_pass(one_pass(r, hp, False), hp, True)
    f = r.reshape(-1)[PAD:PAD + Wn].copy()
    odd_pad(r)
    r = one_pass(one_pass(r, sh, False), sh, True)
    s = r.reshape(-1)[PAD:PAD + Wn]
    got = f + (s - f) * 0.3
    ref = O.apply_k_weighting(xh[None], O.k_weighting_coeffs(48000))[0]
    assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_plan_helpers_and_synth():
    from omega4_b200.batch import synth
    from omega4_b200.batch.partition import stream_block, all_blocks
    x = synth.synth_channel(3, 1, 48000)
    assert x.dtype == np.float32 and np.array_equal(x, synth.synth_channel(3, 1, 48000))      # seeded
    assert not np.array_equal(x, synth.synth_channel(3, 0, 48000))
    assert 0.3 < x.std() < 0.45 and np.abs(x).max() < 1.0
    assert synth.stream_hash(0, 0) != synth.stream_hash(0, 1) != synth.stream_hash(1, 0)
    s = synth.synth_streams(2, 2, 4096, first_stream=5)
    assert np.array_equal(s[1, 0], synth.synth_channel(6, 0, 4096))
    assert all_blocks(10, 4) == [(0, 3), (3, 3), (6, 3), (9, 1)]
    assert all_blocks(2, 4) == [(0, 1), (1, 1), (2, 0), (2, 0)]
    assert sum(c for _, c in all_blocks(8192, 8)) == 8192 and stream_block(8192, 8, 7) == (7168, 1024)
    with pytest.raises(ValueError):
        stream_block(4, 2, 2)


def test_app_post_tables_match_oracle():
    from omega4_b200 import tables
    from oracle import oracle_np as O
    f = np.fft.rfftfreq(2048, 1 / 48000)[:512]
    for ct, vs in (("instrumental", 0.0), ("vocal", 0.5), ("bass_heavy", 0.25)):
        assert np.array_equal(tables.app_compensation_gains(f, ct, vs), O.app_compensation_gains(f, ct, vs))
    bands = tables.mel_band_indices(48000, 2048, 512)
    assert np.array_equal(tables.app_smoothing_factors(bands, 48000, 2048), O.app_smoothing_factors(bands, 48000, 2048))


def test_weighting_programs_match_oracle_sections():
    from omega4_b200 import tables
    from oracle import oracle_np as O
    for sr in (48000, 96000, 44100):
        for mode, ref in (("A", O.a_weighting_sections(sr)), ("C", O.c_weighting_sections(sr))):
            prog = tables.weighting_program(mode, sr)
            assert len(prog["sections"]) == len(ref) and prog["blend"] == 0 and prog["rms_gate"] == 1
            for (b, a), (rb, ra) in zip(prog["sections"], ref):
                assert np.array_equal(b, rb) and np.array_equal(a, ra)
        k = tables.weighting_program("K", sr)
        assert k["blend"] == 1 and len(k["sections"]) == 2
        z = tables.weighting_program("Z", sr)
        assert z["sections"] == [] and z["rms_gate"] == 0
    assert tables.weighting_program("A", 48000)["gain"] == 2.5
