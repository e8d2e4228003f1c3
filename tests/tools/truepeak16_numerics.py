"""numpy emulation of truepeak16_kernel.cuh's arithmetic: a frame pair scaled to unit peak (float32) and rounded to
half, forward transform, delay factors and the three delayed phases by an FFT whose every operation is rounded to
half (radix-2 here, the kernel's radix 16 / 16 / 8 has the same 11 butterfly levels), maxima against the float64
reference (scipy.signal.resample, omega4/panels/professional_meters.py:283-299).  `forward_half=False` emulates the
r02k kernel (float32 forward transform, spectrum rounded to half).

    python tests/tools/truepeak16_numerics.py [SEED] [N_PAIRS]

Test infrastructure (it calls scipy for the reference value); nothing in the product imports it."""
import sys
import numpy as np
import scipy.signal as ss

W = 2048
h = np.float16


def fma_h(a, b, c):
    return (a.astype(np.float32) * b.astype(np.float32) + c.astype(np.float32)).astype(h)


def fft_h(zr, zi):
    """decimation-in-frequency FFT with every operation rounded to half; output order is irrelevant for a maximum"""
    n = len(zr); zr = zr.astype(h); zi = zi.astype(h); m = n
    while m > 1:
        half = m // 2; k = np.arange(half)
        wr = np.cos(-2 * np.pi * k / m).astype(h); wi = np.sin(-2 * np.pi * k / m).astype(h)
        zr = zr.reshape(-1, m); zi = zi.reshape(-1, m)
        ar, ai, br, bi = zr[:, :half], zi[:, :half], zr[:, half:], zi[:, half:]
        sr = (ar + br).astype(h); si = (ai + bi).astype(h); dr = (ar - br).astype(h); di = (ai - bi).astype(h)
        tr = fma_h(dr, np.broadcast_to(wr, dr.shape), -(di * wi).astype(h)); ti = fma_h(dr, np.broadcast_to(wi, dr.shape), (di * wr).astype(h))
        zr = np.concatenate([sr, tr], axis=1).reshape(-1); zi = np.concatenate([si, ti], axis=1).reshape(-1)
        m = half
    return zr, zi


_k = np.arange(W)
_f = np.where(_k < W // 2, _k, _k - W).astype(np.float64)
ROT = []
for _p in (1, 2, 3):
    _R = np.exp(2j * np.pi * _f * _p / (4 * W)); _R[W // 2] = np.cos(np.pi * _p / 4); ROT.append(_R)


def tp_ref(x):
    y = ss.resample(x.astype(np.float64), 4 * len(x)); p = np.max(np.abs(y))
    return 20 * np.log10(p) if p >= 1e-10 else -100.0


_idx = np.arange(W)
_REV = np.zeros(W, dtype=np.int64)
for _b in range(11):
    _REV |= ((_idx >> _b) & 1) << (10 - _b)


def tp_emulated(xa, xb, forward_half=True):
    pk = max(np.abs(xa).max(), 1e-30), max(np.abs(xb).max(), 1e-30)
    z = (xa / pk[0]).astype(np.float32) + 1j * (xb / pk[1]).astype(np.float32)
    if forward_half:
        fr, fi = fft_h(z.real.astype(h), z.imag.astype(h))      # decimation in frequency: bit-reversed order out
        Zr = np.empty(W, dtype=h); Zi = np.empty(W, dtype=h)
        Zr[_REV] = fr; Zi[_REV] = fi
    else:
        Z = np.fft.fft(z.astype(np.complex64)).astype(np.complex64)
        Zr = Z.real.astype(h); Zi = Z.imag.astype(h)
    best = np.array([1.0, 1.0])
    for R in ROT:
        Rr = R.real.astype(h); Ri = R.imag.astype(h)
        sr = fma_h(Zr, Rr, -(Zi * Ri).astype(h)); si = fma_h(Zr, Ri, (Zi * Rr).astype(h))
        yr, yi = fft_h(sr, (-si.astype(np.float32)).astype(h))
        best[0] = max(best[0], np.abs(yr.astype(np.float64)).max() / W); best[1] = max(best[1], np.abs(yi.astype(np.float64)).max() / W)
    return 20 * np.log10(pk[0] * best[0]), 20 * np.log10(pk[1] * best[1])


KINDS = ("sine", "noise", "clipped noise", "square", "two sines", "impulses", "random walk")


def rand_frame(rng):
    win = np.hanning(W); t = np.arange(W) / 48000.
    kind = int(rng.integers(0, 7))
    if kind == 0: x = 0.95 * np.sin(2 * np.pi * rng.uniform(20, 23900) * t + rng.uniform(0, 6.3))
    elif kind == 1: x = rng.standard_normal(W) * 10 ** rng.uniform(-4, -0.5)
    elif kind == 2: x = np.clip(rng.standard_normal(W) * rng.uniform(0.5, 3), -1, 1)
    elif kind == 3: x = np.sign(np.sin(2 * np.pi * rng.uniform(100, 12000) * t + rng.uniform(0, 6))) * 0.9
    elif kind == 4: x = 0.5 * np.sin(2 * np.pi * rng.uniform(20, 23900) * t) + 0.45 * np.sin(2 * np.pi * rng.uniform(20, 23900) * t + 1.0)
    elif kind == 5:
        x = np.zeros(W); x[rng.integers(0, W, 5)] = rng.uniform(-1, 1, 5)
    else:
        x = np.cumsum(rng.standard_normal(W)); x = x / np.abs(x).max() * 0.8
    if rng.random() < 0.7: x = x * win          # the application's frames are Hann windowed, explicit frames need not be
    return kind, x


def run(seed, n_pairs, forward_half=True):
    rng = np.random.default_rng(seed)
    errs, kinds = [], []
    for _ in range(n_pairs):
        ka, xa = rand_frame(rng); kb, xb = rand_frame(rng)
        ea, eb = tp_emulated(xa, xb, forward_half)
        errs += [abs(ea - tp_ref(xa)), abs(eb - tp_ref(xb))]; kinds += [ka, kb]
    return np.array(errs), np.array(kinds)


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 250
    for fh, what in ((True, "all four transforms in half (truepeak16_kernel)"), (False, "float32 forward transform, delayed phases in half (r02k)")):
        e, k = run(seed, n, fh)
        print("%s, %d frames, seed %d: |dBTP error| max %.5f  p99 %.5f  median %.5f  (bar 0.05)" %
              (what, len(e), seed, e.max(), np.percentile(e, 99), np.median(e)))
        for i, name in enumerate(KINDS):
            if (k == i).any(): print("  %-14s max %.5f  median %.5f  (%d frames)" % (name, e[k == i].max(), np.median(e[k == i]), (k == i).sum()))
