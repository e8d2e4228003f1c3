"""world_size=2 test of the stream partitioner + result gather on CPU (gloo).  The per-stream
"analysis" is the numpy oracle on short clips, so the test also shows that sharding streams
across ranks and gathering final rows reproduces the single-process result exactly."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_STREAMS, N_CH, N_HOPS, HOP = 5, 2, 6, 512


def _rows_for_block(first, count):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "audio-analyzer-omega_b200"))
    from omega4_b200.batch.synth import synth_streams
    from oracle import oracle_np as O
    x = synth_streams(count, N_CH, N_HOPS * HOP, first_stream=first)
    rows = np.zeros((count, N_CH, 5))
    for s in range(count):
        for c in range(N_CH):
            rows[s, c] = O.analyze_channel(x[s, c], 48000, O.DEFAULT_CONFIGS)["meters"][-1]
    return rows


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.join(ROOT, "audio-analyzer-omega_b200"))
        from omega4_b200.batch.partition import stream_block, gather_rows
        first, count = stream_block(N_STREAMS, world, rank)
        local = torch.from_numpy(_rows_for_block(first, count))
        full = gather_rows(local, N_STREAMS)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, full.numpy(), float(t[0])))
    finally:
        dist.destroy_process_group()


def test_two_rank_partition_and_gather():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, full, tmax = q.get(timeout=180)
        got[rank] = full
        assert tmax == 2.0
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ref = _rows_for_block(0, N_STREAMS)
    assert got[0].shape == (N_STREAMS, N_CH, 5)
    assert np.array_equal(got[0], ref) and np.array_equal(got[1], ref)


def test_gather_without_process_group_is_identity():
    import torch
    sys.path.insert(0, os.path.join(ROOT, "audio-analyzer-omega_b200"))
    from omega4_b200.batch.partition import gather_rows, final_rows
    x = torch.arange(12.0).reshape(3, 4)
    assert gather_rows(x, 3) is x
    m = torch.arange(2 * 2 * 3 * 5, dtype=torch.float32).reshape(4, 3, 5)
    assert final_rows(m, 2).shape == (2, 2, 5) and torch.equal(final_rows(m, 2)[1, 0], m[2, -1])
