"""Randomised parity stress of the whole hot path against the numpy oracle: streams made of sections with
very different character (digital silence, 1e-7 .. 1e-3 level noise, full-scale tones, clipped bursts, DC
offsets, impulses), through the batch entry point, channel by channel against oracle.analyze_channel.

    python tests/tools/random_stress.py SEED [tc|fft|fp32|tcfd] [config2|config5] [FLOOR_DB]

config2 = BASELINE sizes 8192 .. 1024 at 48 kHz; config5 = 96 kHz, six resolutions 32768 .. 1024 (BASELINE configs[4]).
FLOOR_DB (default -80): the spectrum gate asserts 0.01 dB for every value within that many dB of the largest
magnitude of the transform it came from; the report line also lists the worst dB error at other floors."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "audio-analyzer-omega_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from omega4_b200.plan import AnalysisPlan, BASELINE_CONFIGS, CONFIG5_96K
from oracle import oracle_np as O
from conftest import db

HOP = 512
from omega4_b200 import _native as N
mode = sys.argv[2] if len(sys.argv) > 2 else "tc"
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
cfg_name = sys.argv[3] if len(sys.argv) > 3 else "config2"
FLOOR_DB = float(sys.argv[4]) if len(sys.argv) > 4 else -80.0
if cfg_name == "config5":
    SR, CONFIGS, ORACLE_CONFIGS, n_ch, n_hops = 96000, CONFIG5_96K, O.CONFIG5_96K, 4, 260
else:
    SR, CONFIGS, ORACLE_CONFIGS, n_ch, n_hops = 48000, BASELINE_CONFIGS, O.BASELINE_CONFIGS, 6, 140
FLOORS = (-40.0, -60.0, -80.0, -100.0, -120.0)


def section(kind, n):
    t = np.arange(n) / float(SR)
    if kind == 0: return np.zeros(n)
    if kind == 1: return rng.standard_normal(n) * 10 ** rng.uniform(-7, -3)
    if kind == 2: return 0.95 * np.sin(2 * np.pi * rng.uniform(30, 18000) * t + rng.uniform(0, 6))
    if kind == 3: return np.clip(rng.standard_normal(n) * 2.0, -1, 1)
    if kind == 4: return 0.3 + 0.05 * rng.standard_normal(n)
    if kind == 5:
        x = np.zeros(n); x[rng.integers(0, n, 5)] = rng.uniform(-1, 1, 5); return x
    return rng.standard_normal(n) * 0.1 + 0.4 * np.sin(2 * np.pi * rng.uniform(20, 300) * t)


x = np.zeros((n_ch, n_hops * HOP), np.float32)
for c in range(n_ch):
    pos = 0
    while pos < x.shape[1]:
        n = int(rng.integers(300, 9000 * (SR // 48000) ** 2)); n = min(n, x.shape[1] - pos)
        x[c, pos:pos + n] = section(int(rng.integers(0, 7)), n)
        pos += n
def spectrum_gate(got, ref, mags, label, mode, tol_db=0.01, rel_floor_db=-80.0, by_floor=None):
    """The north star's 0.01 dB wherever the reference value is within |rel_floor_db| dB of the LARGEST magnitude of the
    transform it came from (all N/2+1 bins, DC included: a combined row's own maximum says nothing about the
    float32 noise floor of an 8192-point transform whose window still holds a full-scale burst that the 1024
    window has already left), the same absolute error below that; and the coverage pattern: where the
    reference is exactly zero because no resolution contributes, so is the product.

    The default tensor-core hop-block DFT ("tc": the window is part of the GEMM operand) and the full-FFT mode
    ("fft", OMEGA4_FLAG_NO_BLOCKDFT) are held to this per-frame statement.  The cosine-sum hop-block modes
    ("tcfd" = OMEGA4_BLOCKDFT_FD=1, and "fp32" = OMEGA4_FLAG_NO_TENSOR) apply the window in the frequency
    domain, so their rounding error scales with the UNWINDOWED energy inside the frame: a burst sitting under
    the window's near-zero edge leaves an error of 1e-7 of its own magnitude on a value the window has
    attenuated by 60 dB or more.  For them the 60 dB are counted from the largest magnitude the transform
    shows while the same samples are inside its window (+- N/hop hops)."""
    tf = np.linspace(0.0, 20000.0, ref.shape[1])
    floor = np.zeros(ref.shape)
    peak_tb = np.zeros(ref.shape)                                    # per target bin: the largest magnitude of its transform
    for i, ((lo, hi), n, _h, _w, _t) in enumerate(CONFIGS):
        first, m = mags[i]
        peak = np.zeros(ref.shape[0]); peak[first:first + len(m)] = m.max(axis=1)
        if mode in ("fp32", "tcfd") and n > 2048:                     # resolutions served by the hop-block DFT
            B = n // HOP
            pad = np.pad(peak, B)
            peak = np.max(np.stack([pad[d:d + len(peak)] for d in range(2 * B + 1)]), axis=0)
        sel = (tf >= lo) & (tf <= hi)
        floor[:, sel] = np.maximum(floor[:, sel], peak[:, None] * 10 ** (rel_floor_db / 20.0))
        peak_tb[:, sel] = np.maximum(peak_tb[:, sel], peak[:, None])
    g = got.astype(np.float64); r = ref.astype(np.float64)
    if by_floor is not None:                                         # report only: worst dB error above each floor
        for fl in FLOORS:
            s2 = (r >= np.maximum(peak_tb * 10 ** (fl / 20.0), 1e-30)) & (peak_tb > 0)
            if s2.any():
                by_floor[fl] = max(by_floor.get(fl, 0.0), float(np.abs(db(g[s2], 1e-30) - db(r[s2], 1e-30)).max()))
    uncovered = floor == 0
    assert np.array_equal(g[uncovered] == 0, r[uncovered] == 0), f"{label}: coverage pattern"
    lim = np.maximum(r, floor) * (10 ** (tol_db / 20.0) - 1.0) + 1e-12
    bad = np.abs(g - r) > lim
    assert not bad.any(), f"{label}: {int(bad.sum())} bins off, worst {np.abs(g - r)[bad].max():.3g} at {np.argwhere(bad)[0].tolist()}"
    sig = r >= np.maximum(floor, 1e-30)
    return float(np.abs(db(g[sig], 1e-30) - db(r[sig], 1e-30)).max()) if sig.any() else 0.0


if mode == "tcfd":
    os.environ["OMEGA4_BLOCKDFT_FD"] = "1"
plan = AnalysisPlan(SR, CONFIGS, 512)
got = plan.analyze_host(x, want_series=True, flags={"tc": 0, "tcfd": 0, "fp32": N.FLAG_NO_TENSOR, "fft": N.FLAG_NO_BLOCKDFT}[mode])
worst = {"spec_db": 0.0, "lufs": 0.0, "tp": 0.0, "meters": 0.0}
by_floor = {}
for c in range(n_ch):
    ref = O.analyze_channel(x[c], SR, ORACLE_CONFIGS, keep_magnitudes=True)
    worst["spec_db"] = max(worst["spec_db"], spectrum_gate(got["combined"][c], ref["combined"], ref["magnitudes"], f"ch {c}", mode,
                                                           rel_floor_db=FLOOR_DB, by_floor=by_floor))
    m = ~np.isnan(ref["lufs_inst"])
    dl = np.abs(got["lufs_inst"][c][m] - ref["lufs_inst"][m]); dt = np.abs(got["tp_db"][c][m] - ref["tp_db"][m])
    dm = np.abs(got["meters"][c] - ref["meters"])
    worst["lufs"] = max(worst["lufs"], dl.max()); worst["tp"] = max(worst["tp"], dt.max())
    worst["meters"] = max(worst["meters"], dm[:, :4].max())
    if dl.max() > 0.01 or dt.max() > 0.05 or dm[:, :4].max() > 0.01 or dm[:, 4].max() > 0.05:
        k = int(np.argmax(dt)); print("FAIL ch", c, "lufs", dl.max(), "tp", dt.max(), "at frame", np.flatnonzero(m)[k],
                                      got["tp_db"][c][m][k], ref["tp_db"][m][k], "meters", dm.max(axis=0))
        sys.exit(1)
print(f"random stress ok ({mode}, {cfg_name}, gate floor {FLOOR_DB:g} dB):", {k: float(f"{v:.3g}") for k, v in worst.items()},
      "| worst dB err above floor:", {f"{int(k)}": float(f"{v:.2g}") for k, v in by_floor.items()})
