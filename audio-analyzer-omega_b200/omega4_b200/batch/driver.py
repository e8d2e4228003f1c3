"""Headless batch driver: feeds multi-stream audio through the GPU hot path, decoupled from pygame.

Replaces the caller side of the reference for this path -- ``process_audio_spectrum`` /
``process_multi_resolution_fft`` (omega4_main.py:690-778, 928-1082) and the in-module
``benchmark_multi_fft`` driver (omega4/audio/multi_resolution_fft.py:467-494) -- on the shared
frame schedule of SURVEY.md section 7 step 2: hop 512; resolution N contributes to hop k once
(k+1)*512 >= N; one meter update per hop on the Hann-windowed last 2048 samples.

Which feed this is: the batch schedule hands every resolution the last N RAW samples at each 512-sample
hop -- the feed of the reference's own headless driver ``benchmark_multi_fft`` (512-sample chunks into
the rings).  The pygame application feeds differently: it passes the Hann-windowed 2048-sample window to
``process_audio_chunk`` once per video frame (omega4_main.py:953-980), so its rings hold overlapping,
pre-windowed blocks; that feed is reproduced by the streaming shim
(``omega4_b200.audio.multi_resolution_fft.MultiResolutionFFT``, golden ``multires_appfeed.npz``), not here.

``StreamBatch`` keeps the per-channel carry between time tiles (the last max(N)-hop samples and
the meters' deque state), so arbitrarily long streams can be processed tile by tile -- e.g. when
the audio is generated on the device because it does not fit in host memory (BASELINE config 4).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .. import _native as N
from ..plan import AnalysisPlan


def device_synth(n_streams: int, n_channels: int, n_samples: int, sample_rate: int = 48000,
                 first_stream: int = 0, clip_samples: Optional[int] = None, device: int = 0, out=None):
    """float32 CUDA tensor [n_streams*n_channels, n_samples] of the benchmark signal, generated
    on the device by ``omega4_synth_fill`` (sweep + white counter-hash noise)."""
    import torch
    rows = n_streams * n_channels
    if out is None:
        out = torch.empty((rows, n_samples), dtype=torch.float32, device=f"cuda:{device}")
    assert out.is_cuda and out.dtype == torch.float32 and out.shape[0] == rows and out.stride(1) == 1
    stream = torch.cuda.current_stream(out.device).cuda_stream
    done = 0
    while done < n_streams:                      # gridDim.y limit: 65535 rows per launch
        cnt = min(n_streams - done, 65535 // n_channels)
        rc = N.lib().omega4_synth_fill(device, stream, out.data_ptr() + done * n_channels * out.stride(0) * 4,
                                       cnt, n_channels, n_samples, out.stride(0), first_stream + done, sample_rate,
                                       clip_samples or n_samples)
        N.check(rc, "omega4_synth_fill")
        done += cnt
    return out


class StreamBatch:
    """A batch of independent channels processed tile by tile on one GPU."""

    def __init__(self, plan: AnalysisPlan, n_channels_total: int, max_tile_hops: int):
        import torch
        self.plan = plan
        self.n_ch = int(n_channels_total)
        self.max_tile_hops = int(max_tile_hops)
        self.dev = torch.device(f"cuda:{plan.device}")
        self.hist_cap = max(max(plan.sizes), plan.meter_window) - plan.hop
        self.hist_cap = (self.hist_cap + 3) // 4 * 4
        self.row = self.hist_cap + self.max_tile_hops * plan.hop
        self.buf = torch.zeros((self.n_ch, self.row), dtype=torch.float32, device=self.dev)
        self.state = torch.zeros((self.n_ch, N.METER_STATE_DOUBLES), dtype=torch.float64, device=self.dev)
        self.hist = 0                 # valid history samples currently in buf[:, hist_cap-hist:hist_cap]
        self.hops_done = 0

    def tile_view(self, n_hops: int):
        """Where the caller writes the next tile's samples: float32 [n_ch, n_hops*hop] view."""
        return self.buf[:, self.hist_cap:self.hist_cap + n_hops * self.plan.hop]

    def push(self, n_hops: int, combined=None, meters=None, magnitudes=None, flags: int = 0):
        """Analyse the tile previously written into ``tile_view(n_hops)``."""
        assert 0 < n_hops <= self.max_tile_hops
        fl = flags | (N.FLAG_FRESH_METERS if self.hops_done == 0 else 0)
        self.plan.analyze_device(self.buf[:, self.hist_cap - self.hist:], n_hops, hist_samples=self.hist,
                                 combined=combined, magnitudes=magnitudes, meters=meters,
                                 meter_state=self.state, flags=fl)
        # carry: the last hist_cap samples become the history of the next tile
        new_total = self.hist + n_hops * self.plan.hop
        keep = min(self.hist_cap, new_total)
        end = self.hist_cap + n_hops * self.plan.hop
        src = self.buf[:, end - keep:end]
        if end - keep < self.hist_cap:            # overlapping ranges -> go through a temporary
            src = src.clone()
        self.buf[:, self.hist_cap - keep:self.hist_cap].copy_(src)
        self.hist = keep
        self.hops_done += n_hops


def analyze_resident(plan: AnalysisPlan, samples, combined=None, meters=None, magnitudes=None, flags: int = 0):
    """One call over HBM-resident audio: ``samples`` float32 CUDA tensor [n_ch, n_hops*hop (+tail)]."""
    n_hops = samples.shape[1] // plan.hop
    plan.analyze_device(samples, n_hops, 0, combined=combined, magnitudes=magnitudes, meters=meters,
                        flags=flags | N.FLAG_FRESH_METERS)
    return n_hops


def analyze_numpy(plan: AnalysisPlan, samples: np.ndarray, **kw) -> Dict[str, object]:
    """Convenience: host arrays in and out (see AnalysisPlan.analyze_host)."""
    return plan.analyze_host(samples, **kw)
