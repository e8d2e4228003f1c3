"""CPU check of the half-precision true-peak arithmetic (truepeak16_kernel.cuh) through its numpy emulation:
the emulated peak stays well inside the north star's 0.05 dBTP of scipy.signal.resample's."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))


def test_half_precision_true_peak_emulation_is_inside_the_gate():
    import truepeak16_numerics as T
    errs, _ = T.run(3, 30)
    assert errs.max() < 0.03, errs.max()
    assert sorted(errs)[len(errs) // 2] < 0.006
