"""Stream partitioner for multi-GPU runs.

Streams (and channels) are fully independent, so the path shards with NO data-path collective:
rank r of G owns a contiguous block of ceil(S/G) streams, generates / loads and analyses only
those, and the only communication is ONE gather of the final per-stream result rows after the
last kernel -- a few bytes per stream, never on the hot path.  ``torch.distributed`` provides the
plumbing: NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def stream_block(n_streams: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(first_stream, count) of the contiguous block owned by ``rank``."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    per = (n_streams + world_size - 1) // world_size
    first = min(rank * per, n_streams)
    return first, max(0, min(per, n_streams - first))


def all_blocks(n_streams: int, world_size: int) -> List[Tuple[int, int]]:
    return [stream_block(n_streams, world_size, r) for r in range(world_size)]


def gather_rows(local_rows, n_streams: int, group=None):
    """All-gather per-stream result rows.  ``local_rows``: tensor [count_r, ...] for this rank's
    block (same trailing shape on every rank).  Returns tensor [n_streams, ...] on every rank, in
    global stream order.  Blocks are padded to the common ceil(S/G) length for the collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_rows
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_streams + world - 1) // world
    first, count = stream_block(n_streams, world, rank)
    assert local_rows.shape[0] == count, (local_rows.shape, count)
    pad = torch.zeros((per,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    pad[:count] = local_rows
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    parts = []
    for r in range(world):
        _, c = stream_block(n_streams, world, r)
        parts.append(out[r][:c])
    return torch.cat(parts, dim=0)


def final_rows(meters, n_channels: int):
    """Per-stream summary row from a meters tensor [n_ch_total, n_hops, 5]: the last hop's
    (M, S, I, LRA, TP) of every channel -> [n_streams, n_channels, 5]."""
    last = meters[:, -1, :]
    return last.reshape(-1, n_channels, last.shape[-1])
