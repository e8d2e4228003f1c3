"""Pinned host <-> device copy bandwidth of this box: one direction, both at once, and -- under torchrun -- on
every GPU at the same time, which is the ceiling of bench.py's end-to-end leg (its result rows and samples all
cross the bus; nothing else in that leg touches host memory).

    python tools/pcie_bw.py                                   one GPU
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_bw.py [--no-numa]

Prints one JSON line (rank 0): per-GPU minimum / mean and the aggregate GB/s over all ranks, measured over the
same wall-clock window (barrier on both sides), with the ranks bound to their GPU's NUMA node the way bench.py
binds them (or not, with --no-numa)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import bind_to_gpu_numa_node  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
numa = {"how": "disabled"} if "--no-numa" in sys.argv else bind_to_gpu_numa_node(local, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")

n = 1 << 28                      # 1 GiB of float32 per buffer
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
h_in.fill_(1.0); h_out.fill_(0.0)                 # first touch after the binding
d_a = torch.empty(n, dtype=torch.float32, device="cuda")
d_b = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = n * 4 / 1e9


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timeit(fn, reps=6):
    fn(); barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    barrier()
    return (time.perf_counter() - t0) / reps


def h2d():
    d_a.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_b, non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def e2e_mix():
    """bench.py's e2e traffic mix: int16 samples up (2 B per sample), float32 result rows down (4 B per sample's
    worth of spectrum): 1 byte up for every 2 bytes down, both directions busy at once."""
    with torch.cuda.stream(s1):
        d_a[: n // 2].copy_(h_in[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


res = {}
for name, fn, nbytes in (("h2d", h2d, gb), ("d2h", d2h, gb), ("duplex_each_direction", both, gb), ("e2e_mix_total", e2e_mix, 1.5 * gb)):
    t = timeit(fn)                                 # every rank measures the same window: max over ranks = common time
    res[name] = nbytes / t
if world > 1:
    vals = torch.tensor([res["h2d"], res["d2h"], res["duplex_each_direction"], res["e2e_mix_total"]], dtype=torch.float64)
    allv = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
    allv = torch.stack(allv)
else:
    allv = torch.tensor([[res["h2d"], res["d2h"], res["duplex_each_direction"], res["e2e_mix_total"]]], dtype=torch.float64)
if rank == 0:
    out = {"gpus": world, "numa": numa, "buffer_gib": 1}
    for j, name in enumerate(("h2d", "d2h", "duplex_each_direction", "e2e_mix_total")):
        col = allv[:, j]
        out[name] = {"per_gpu_min_gbs": float(col.min()), "per_gpu_mean_gbs": float(col.mean()), "aggregate_gbs": float(col.sum())}
    out["duplex_bus_total_gbs"] = 2 * out["duplex_each_direction"]["aggregate_gbs"]
    out["note"] = "e2e_mix_total = bytes in both directions per second with 1 byte up per 2 bytes down (bench.py's e2e leg): its bus_gbs ceiling"
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
