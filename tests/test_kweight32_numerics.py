"""CPU check of the float32-state K-weighting design (kweight32_kernel.cuh): the numpy emulation of the kernel's
lane / sub-chunk / scan structure in float32 stays within 1e-4 LU of the float64 oracle (bar: 0.01 LU) on
Hann-windowed frames of noise, near-cut-off tones, DC offsets and steps."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))


import pytest


@pytest.mark.parametrize("sr", [48000, 96000])
def test_float32_delta_form_matches_float64_oracle(sr):
    """96 kHz (BASELINE configs[4]) moves the 38 Hz section's double pole to radius 0.99824 and shrinks
    alpha four-fold; the delta form keeps its margin there (profiles/r02_kweight32_numerics_96k.txt)."""
    import kweight32_numerics as K
    from oracle import oracle_np as O
    rng = np.random.default_rng(7)
    n = 2048
    t = np.arange(n) / float(sr)
    hann = np.hanning(n)
    frames = np.concatenate([
        rng.standard_normal((8, n)) * 0.1,
        rng.standard_normal((4, n)) * 1e-4,
        0.3 + rng.standard_normal((4, n)) * 1e-4,
        np.stack([0.9 * np.sin(2 * np.pi * f0 * t + 1.0) for f0 in (10.0, 30.0, 38.0, 1000.0, 10000.0)]),
        np.where(np.arange(n)[None, :] >= np.array([[300], [1500]]), 0.8, 0.0),
    ])
    x = (frames.astype(np.float32).astype(np.float64) * hann).astype(np.float32).astype(np.float64)
    ref = O.lufs_instantaneous(x, O.k_weighting_coeffs(sr))
    got = K.lufs32(x, sr)
    assert np.abs(got - ref).max() < 1e-4, np.abs(got - ref).max()
