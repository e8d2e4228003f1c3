"""Randomised adversarial parity of the whole hot path (tests/tools/random_stress.py): streams cut from sections of
digital silence, 1e-7 .. 1e-3 noise, full-scale tones, clipped bursts, DC offsets and isolated impulses,
through the C ABI's batch entry point against the numpy oracle, channel by channel.  The tool states the
spectrum gate it applies to each mode (the default path and the full FFT: per frame; the cosine-sum hop-block
modes count the 60 dB from the transform's largest magnitude while the same samples are inside its window)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed,mode", [(1, "tc"), (3, "fp32"), (6, "fft"), (10, "tc"), (5, "tcfd")])
def test_random_sections_against_oracle(seed, mode):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "random_stress.py"), str(seed), mode],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "random stress ok" in out.stdout
