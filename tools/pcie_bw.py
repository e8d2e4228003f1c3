"""Pinned host <-> device copy bandwidth on this box (one direction and both at once)."""
import time
import torch

n = 1 << 28                      # 1 GiB of float32
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device="cuda")
d_b = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = n * 4 / 1e9


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


t = timeit(lambda: d_a.copy_(h_in, non_blocking=True))
print(f"H2D pinned  : {gb / t:6.1f} GB/s")
t = timeit(lambda: h_out.copy_(d_b, non_blocking=True))
print(f"D2H pinned  : {gb / t:6.1f} GB/s")


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


t = timeit(both)
print(f"H2D + D2H concurrently: {gb / t:6.1f} GB/s each direction")
