"""Seeded synthetic multi-stream audio for the headless batch driver (host side, numpy).

The measurement plan (SURVEY.md section 8d / BASELINE.md section 3) feeds every arm -- oracle,
CPU baseline and GPU -- the same signal: per (stream, channel) a log sweep 20 Hz -> 20 kHz over
the clip, amplitude 0.5, start phase derived from an integer hash of (stream, channel), plus
pink noise (white ``default_rng(1000*stream + channel)`` shaped by 1/sqrt(f) in the rFFT
domain) at RMS 0.1; float32.

The device-side generator used by ``bench.py`` for the HBM-resident workload lives in
``csrc/omega4_cuda.cu`` (``omega4_synth_fill``): same sweep, counter-hash white noise.
"""
from __future__ import annotations

import numpy as np

SWEEP_F0 = 20.0
SWEEP_F1 = 20000.0
SWEEP_AMP = 0.5
NOISE_RMS = 0.1


def stream_hash(stream: int, channel: int) -> int:
    """32-bit integer hash of (stream, channel); also implemented in omega4_synth_fill."""
    h = (stream * 0x9E3779B1 + channel * 0x85EBCA77 + 0x165667B1) & 0xFFFFFFFF
    h ^= h >> 15
    h = (h * 0x2C1B3C6D) & 0xFFFFFFFF
    h ^= h >> 12
    h = (h * 0x297A2D39) & 0xFFFFFFFF
    h ^= h >> 15
    return h


def sweep(n_samples: int, sample_rate: int, stream: int = 0, channel: int = 0,
          clip_samples: int | None = None) -> np.ndarray:
    """0.5*sin(2 pi f0 (e^{kt}-1)/k + phi0), k = ln(f1/f0)/T, T = clip duration; float64."""
    clip = n_samples if clip_samples is None else clip_samples
    T = clip / sample_rate
    k = np.log(SWEEP_F1 / SWEEP_F0) / T
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    phi0 = 2.0 * np.pi * stream_hash(stream, channel) / 2.0 ** 32
    return SWEEP_AMP * np.sin(2.0 * np.pi * SWEEP_F0 * np.expm1(k * t) / k + phi0)


def pink_noise(n_samples: int, seed: int) -> np.ndarray:
    """White gaussian noise shaped by 1/sqrt(f) in the rFFT domain, normalised to RMS 0.1."""
    rng = np.random.default_rng(seed)
    white = rng.standard_normal(n_samples)
    X = np.fft.rfft(white)
    f = np.arange(len(X), dtype=np.float64)
    f[0] = 1.0
    X /= np.sqrt(f)
    X[0] = 0.0
    pink = np.fft.irfft(X, n_samples)
    rms = np.sqrt(np.mean(pink ** 2))
    return pink * (NOISE_RMS / rms) if rms > 0 else pink


def synth_channel(stream: int, channel: int, n_samples: int, sample_rate: int = 48000,
                  clip_samples: int | None = None) -> np.ndarray:
    """One channel of the benchmark signal, float32[n_samples]."""
    x = sweep(n_samples, sample_rate, stream, channel, clip_samples)
    x = x + pink_noise(n_samples, 1000 * stream + channel)
    return x.astype(np.float32)


def synth_streams(n_streams: int, n_channels: int, n_samples: int, sample_rate: int = 48000,
                  first_stream: int = 0) -> np.ndarray:
    """float32[n_streams, n_channels, n_samples]; stream ids start at ``first_stream`` so a
    partitioned run generates exactly the rows a single-process run would."""
    out = np.empty((n_streams, n_channels, n_samples), dtype=np.float32)
    for s in range(n_streams):
        for c in range(n_channels):
            out[s, c] = synth_channel(first_stream + s, c, n_samples, sample_rate)
    return out
