"""Numerics study behind kweight32_kernel.cuh (checker-side: lives under tests/ because it uses the oracle).

A numpy float32 emulation of the kernel's exact structure -- 32 lanes x 65 samples, four zero-state sub-chunk
sweeps per lane, Kogge-Stone scan of the (z, d) chunk states, homogeneous correction, forward and backward
pass, two high-pass sections, blend -- against the float64 oracle's instantaneous LUFS, on Hann-windowed frames
(what the batch path feeds) and on raw frames:

    python tests/tools/kweight32_numerics.py [sample_rate]

Every float32 operation is rounded separately here (numpy has no FMA), so the CUDA kernel is slightly better.
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle_np as O
f32 = np.float32
L, W, PAD = 65, 2048, 9
EXT = W + 2*PAD
OFF = [0, 17, 33, 49, 65]

def consts(b, a):
    """Kw32Sec as build_kw32 derives it (csrc/omega4_cuda.cu): everything in float64, then rounded once."""
    b0 = b[0]; a1, a2 = a[1], a[2]
    alpha = 1.0 + a1 + a2; beta = 1.0 - a2
    bb = -b0 * beta; gamma = alpha / bb
    M = np.array([[1 - alpha, bb * a2], [-gamma, a2]])
    def P(k): return np.linalg.matrix_power(M, k)
    c = dict(b0=f32(b0), a2=f32(a2), gamma=f32(gamma), bb=f32(bb))
    c['phi'] = [P(65 * 2**j).astype(f32) for j in range(5)]
    c['c16'] = P(16).astype(f32); c['c17'] = P(17).astype(f32)
    c['g'] = np.stack([P(i + 1)[0] for i in range(17)]).astype(f32)
    return c

def sweep_pass(r, c, backward):
    """kw32_pass: y = b0 x + z;  delta = a2 delta' + (x' - x'') - gamma y';  z = z' + bb delta.
    Zero-state sub-chunk sweeps + scan of the (z, delta) chunk states + homogeneous correction, float32."""
    F = r.shape[0]
    if backward:
        r = r[:, ::-1, ::-1]
    xa = np.empty((F, 32), f32); xb = np.empty((F, 32), f32)
    xa[:, 1:] = r[:, :-1, 64]; xb[:, 1:] = r[:, :-1, 63]
    xa[:, 0] = r[:, 0, 0]; xb[:, 0] = r[:, 0, 0]
    offs = OFF if not backward else [0, 16, 32, 48, 65]
    rin = r.copy()
    ys = np.empty_like(rin)
    esub = []
    for m in range(4):
        lo, hi = offs[m], offs[m + 1]
        if m == 0: pa, dx = xa, xa - xb
        else: pa, dx = rin[:, :, lo - 1], rin[:, :, lo - 1] - rin[:, :, lo - 2]
        z1 = np.zeros((F, 32), f32); d1 = np.zeros((F, 32), f32); y1 = c['b0'] * pa
        for i in range(lo, hi):
            x = rin[:, :, i]
            t = c['a2'] * d1 + dx
            d = t - c['gamma'] * y1
            z = z1 + c['bb'] * d
            y = c['b0'] * x + z
            dx = x - pa; pa = x; z1 = z; d1 = d; y1 = y
            ys[:, :, i] = y
        esub.append(np.stack((z1, d1), -1))
    v = esub[0]
    for m in range(1, 4):
        cm = c['c17'] if (offs[m + 1] - offs[m]) == 17 else c['c16']
        v = (v @ cm.T).astype(f32) + esub[m]
    s0 = np.zeros((F, 2), f32); s0[:, 0] = -c['b0'] * xa[:, 0]         # y = 0 in steady state: z = -b0 x0
    v[:, 0] = v[:, 0] + (s0 @ c['phi'][0].T).astype(f32)
    for j in range(5):
        dd = 1 << j
        sh = np.zeros_like(v); sh[:, dd:] = v[:, :-dd]
        add = (sh @ c['phi'][j].T).astype(f32); add[:, :dd] = 0
        v = v + add
    sin = np.zeros_like(v); sin[:, 1:] = v[:, :-1]; sin[:, 0] = s0
    for m in range(4):
        lo, hi = offs[m], offs[m + 1]
        if m > 0:
            cm = c['c17'] if (offs[m] - offs[m - 1]) == 17 else c['c16']
            sin = (sin @ cm.T).astype(f32) + esub[m - 1]
        for n, i in enumerate(range(lo, hi)):
            r[:, :, i] = ys[:, :, i] + (c['g'][n, 0] * sin[:, :, 0] + c['g'][n, 1] * sin[:, :, 1])

def filtfilt32(x, c):
    F = x.shape[0]
    r = np.zeros((F, 32 * L), f32)
    r[:, PAD:PAD + W] = x
    r[:, :PAD] = 2 * x[:, :1] - x[:, PAD:0:-1]
    r[:, PAD + W:EXT] = 2 * x[:, -1:] - x[:, -2:-(PAD + 2):-1]
    r[:, EXT:] = r[:, EXT - 1:EXT]
    r = r.reshape(F, 32, L)
    sweep_pass(r, c, False)
    rr = r.reshape(F, -1); rr[:, EXT:] = rr[:, EXT - 1:EXT]
    sweep_pass(r, c, True)
    return r.reshape(F, -1)[:, PAD:PAD + W].copy()

SR = int(sys.argv[1]) if __name__ == '__main__' and len(sys.argv) > 1 else 48000


def lufs32(frames64, sample_rate=None):
    co = O.k_weighting_coeffs(sample_rate or SR)
    c1 = consts(co['hp_b'], co['hp_a']); c2 = consts(co['shelf_b'], co['shelf_a'])
    x = frames64.astype(f32)
    f = filtfilt32(x, c1)
    s = filtfilt32(f, c2)
    w = f + (s - f) * f32(0.3)
    ms = (w.astype(np.float64) ** 2).mean(-1)
    with np.errstate(divide='ignore'):
        l = -0.691 + 10 * np.log10(ms)
    return np.where(ms > 1e-10, l, -100.0)


if __name__ == '__main__':
    rng = np.random.default_rng(0)
    n = 2048; t = np.arange(n) / float(SR)
    hann = np.hanning(n)
    cases = {}
    cases['white 0.1'] = rng.standard_normal((64, n)) * 0.1
    cases['white 1e-4'] = rng.standard_normal((64, n)) * 1e-4
    cases['dc0.3+noise1e-4'] = 0.3 + rng.standard_normal((64, n)) * 1e-4
    cases['dc0.9'] = np.full((8, n), 0.9)
    for f0 in (10, 20, 30, 38, 50, 100, 1000, 10000):
        cases[f'tone {f0}'] = 0.9 * np.sin(2 * np.pi * f0 * t[None, :] + rng.uniform(0, 6, (16, 1)))
    cases['tone30+tiny hf'] = 0.9 * np.sin(2 * np.pi * 30 * t)[None, :] + 1e-4 * rng.standard_normal((16, n))
    cases['clipped'] = np.clip(rng.standard_normal((64, n)) * 2, -1, 1)
    imp = np.zeros((32, n)); imp[np.arange(32), rng.integers(0, n, 32)] = 1.0; cases['impulse'] = imp
    step = np.zeros((32, n)); 
    for i in range(32): step[i, rng.integers(100, n-100):] = 0.8
    cases['step'] = step
    co = O.k_weighting_coeffs(SR)
    for windowed in (True, False):
        print('hann windowed' if windowed else 'raw frames')
        for k, fr in cases.items():
            fr32 = fr.astype(f32).astype(np.float64)
            xin = (fr32 * hann).astype(f32).astype(np.float64) if windowed else fr32
            ref = O.lufs_instantaneous(xin, co)
            got = lufs32(xin)
            print(f'  {k:20s} ref {ref.mean():9.3f}  max |dLUFS| {np.abs(got - ref).max():.3e}')
