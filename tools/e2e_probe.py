"""Developer probe: host-buffer (e2e) throughput of omega4_analyze vs chunk size / streams per call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "audio-analyzer-omega_b200"))
import torch
from omega4_b200 import _native as N
from omega4_b200.plan import AnalysisPlan, BASELINE_CONFIGS
from omega4_b200.batch.driver import device_synth

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 128
secs = 60
n_hops = secs * 48000 // 512
n_samples = n_hops * 512
ech = streams * 2
x = device_synth(streams, 2, n_samples)
hx = torch.empty((ech, n_samples), dtype=torch.float32, pin_memory=True); hx.copy_(x); del x
hcomb = torch.empty((ech, n_hops, 512), dtype=torch.float32, pin_memory=True)
hmet = torch.empty((ech, n_hops, 5), dtype=torch.float32, pin_memory=True)
for mb in (int(v) for v in os.environ.get("MBS", "256,768,2048").split(",")):
    os.environ["OMEGA4_HOST_CHUNK_MB"] = str(mb)
    plan = AnalysisPlan(48000, BASELINE_CONFIGS, 512)
    def call():
        rc = N.lib().omega4_analyze(plan.handle, None, N.MEM_HOST, hx.data_ptr(), n_samples, ech, n_hops, 0,
                                    hcomb.data_ptr(), None, hmet.data_ptr(), None, None, None, 0)
        N.check(rc)
    call(); torch.cuda.synchronize()
    t0 = time.perf_counter(); reps = 3
    for _ in range(reps): call()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    gb = (hx.numel() + hcomb.numel() + hmet.numel()) * 4 / 1e9
    print(f"streams/call {streams} chunk {mb} MB: {dt*1e3:.1f} ms/call -> {streams*secs/dt:.0f} stream-s/s, {gb/dt:.1f} GB/s total PCIe", flush=True)
    plan.close()
