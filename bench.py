#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): stream-seconds of 48 kHz stereo analysed per second through the full
multi-resolution FFT (8192/4096/2048/1024, hop 512) + combine + LUFS (M/S/I/LRA) + 4x true peak
pipeline.  Workload = BASELINE config[2]: 1024 concurrent stereo streams x 60 s per GPU (weak
scaling: every rank analyses its own 1024 streams -- streams are independent, no data-path
collective; config[3]'s 8192 streams / 8 GPUs is the same per-rank shape).

One step = one pass of the whole path over that batch.  Printed JSON line (rank 0):
  value         device-resident throughput (inputs already in HBM), CUDA events, max over ranks
  e2e           same metric through the C ABI with HOST (pinned) buffers: H2D of every sample and
                D2H of every result inside the timed region
  roofline      dominant kernel vs the measured HBM peak (this path is fp32-issue bound: see
                DESIGN.md section 5; roofline_fp32 gives the compute-side fraction)
  cpu_baseline  the reference's numpy/scipy CPU path (oracle port) on this box's host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "audio-analyzer-omega_b200"))

METRIC = "stream-seconds of 48 kHz stereo analysed/sec (multi-res FFT+LUFS+TP)"
UNIT = "stream-s/s"
SR, HOP, T_BINS, CHANNELS = 48000, 512, 512, 2
B_ALG_PER_HOP = 2048 + 2048 + 20          # SURVEY.md section 8d, fused-output mode: 512 f32 in, 512 f32 + 5 f32 out
FLOP_PER_HOP = 0.9e6                      # SURVEY.md section 8d estimate (fp32 + the fp64 biquads)
FP32_PEAK_TFLOPS = 70.8                   # measured FFMA peak on this pool's B200 (profiles/r01_fp_pipes_microbench.txt:
                                          # 121.7 lanes/clk/SM x 148 SMs x 1.965 GHz x 2); nominal 75


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(sample_seconds, streams=None, repeat=1):
    """Times oracle/ref_port.py (the reference's per-hop numpy/scipy path) in a fresh interpreter
    on all host cores; returns (list of results, cores)."""
    cores = os.cpu_count() or 1
    n = streams or cores
    cmd = [sys.executable, "-m", "oracle.ref_port", "--streams", str(n), "--seconds", str(sample_seconds),
           "--processes", str(cores), "--repeat", str(repeat)]
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", PYTHONDONTWRITEBYTECODE="1")
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    if out.returncode != 0:
        raise RuntimeError("cpu baseline failed: " + out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1]), cores, n


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: the
    reference is Python and is not on the GPU box) on the host cores, same metric and config."""
    if rank != 0:
        return
    secs = args.cpu_seconds
    res, cores, n = cpu_baseline(secs, repeat=args.warmup + args.steps)
    timed = res[args.warmup:]
    wall = sum(r["wall_s"] for r in timed)
    value = sum(r["stream_seconds"] for r in timed) / wall
    sample = f"{n} stereo streams x {secs:g} s per step, {cores} processes (OMP_NUM_THREADS=1)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, len(timed)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 FFT / f64 meters (numpy, scipy)", "data": "synthetic",
        "config": {"workload": "BASELINE config[2] shape: stereo 48 kHz streams, 4 resolutions 8192/4096/2048/1024, hop 512, "
                               "512 target bins, LUFS M/S/I/LRA + 4x true peak; bounded sample per step", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="stereo streams per GPU (config[2]: 1024)")
    ap.add_argument("--seconds", type=int, default=60, help="seconds per stream (config[2]: 60)")
    ap.add_argument("--e2e-streams", type=int, default=1024,
                    help="streams per host-buffer call of the e2e leg (halved until the pinned buffers can be allocated)")
    ap.add_argument("--cpu-seconds", type=float, default=60.0,
                    help="seconds per stream of the CPU baseline sample (one stream per host core; 60 s = the workload's stream length)")
    ap.add_argument("--workload", default="config2", choices=["config2", "config5"],
                    help="config2 (default, the metric's configuration): 48 kHz stereo, 8192/4096/2048/1024; "
                         "config5: 96 kHz 8-channel (7.1) streams, six resolutions 32768 .. 1024 (BASELINE configs[4] shape; "
                         "an extra workload, reported with its own config.workload string)")
    ap.add_argument("--flags", type=int, default=0, help="extra omega4_analyze flags (developer A/B runs, e.g. 64 = OMEGA4_FLAG_SERIAL_STATS)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    from omega4_b200 import _native as N
    from omega4_b200.plan import AnalysisPlan, BASELINE_CONFIGS, CONFIG5_96K
    from omega4_b200.batch.driver import device_synth
    from omega4_b200.batch.partition import gather_rows, final_rows

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    global SR, CHANNELS
    configs, wl_name = BASELINE_CONFIGS, "BASELINE config[2]"
    if args.workload == "config5":
        SR, CHANNELS, configs, wl_name = 96000, 8, CONFIG5_96K, "BASELINE configs[4] shape (96 kHz, 8 channels, 6 resolutions)"
        args.no_cpu = True                       # the CPU leg times the metric's configuration only
    n_streams, n_ch = args.streams, args.streams * CHANNELS
    n_hops = args.seconds * SR // HOP
    n_samples = n_hops * HOP
    plan = AnalysisPlan(SR, configs, T_BINS, device=local)
    first_stream = rank * n_streams
    x = device_synth(n_streams, CHANNELS, n_samples, SR, first_stream=first_stream, device=local)
    comb = torch.empty((n_ch, n_hops, T_BINS), dtype=torch.float32, device=dev)
    met = torch.empty((n_ch, n_hops, N.N_METERS), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(flags=0):
        plan.analyze_device(x, n_hops, 0, combined=comb, meters=met, flags=flags | args.flags | N.FLAG_FRESH_METERS)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = plan.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ktimes = {}
    ev0.record()
    for _ in range(args.steps):
        step(N.FLAG_TIME_KERNELS)            # the library brackets each kernel with events on this stream
    ev1.record()
    barrier()
    # kernel times of the last timed step (events of earlier steps are overwritten; same launches)
    for name, ms in plan.kernel_times():
        ktimes[name] = ms
    launches = plan.launches - launches0

    t_ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(t_ms[0])
    ms_per_step = total_ms / args.steps
    stream_seconds_per_step = n_streams * args.seconds * world
    value = stream_seconds_per_step / (ms_per_step / 1e3)

    # the only collective of the whole path: gather the final per-stream rows (after the timed region)
    rows = final_rows(met, CHANNELS)
    allrows = gather_rows(rows, n_streams * world) if world > 1 else rows
    checksum = float(allrows.double().sum())

    # ---------------- SURVEY section 8f rows, timed separately (not part of the headline metric)
    next_rows = {}
    try:
        from omega4_b200.app.spectrum_post import SpectrumPostProcessor
        if args.workload != "config2":
            raise RuntimeError("timed for the metric's configuration only")
        post = SpectrumPostProcessor(T_BINS, SR, 2048, device=local)
        post._ensure()
        bars = torch.empty((n_ch, n_hops, post.n_valid), dtype=torch.float32, device=dev)
        post.process_device(comb, bars)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        post.process_device(comb, bars)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        next_rows["app_post_processing"] = {"ms": ms, "bars_per_frame": post.n_valid,
                                            "gbs": (comb.numel() + bars.numel()) * 4 / (ms / 1e3) / 1e9,
                                            "frac_hbm": (comb.numel() + bars.numel()) * 4 / (ms / 1e3) / 1e9 / float(measured_peaks()[0].get("hbm_gbs", 6650.0))}
        del bars
        post.close()
    except Exception as e:                       # never let an auxiliary row break the headline line
        next_rows["app_post_processing"] = {"error": str(e)}

    # ---------------- e2e: host (pinned) buffers through the C ABI
    e2e = None
    if not args.no_e2e:
        es = min(args.e2e_streams, n_streams)
        try:                                          # never pin more than ~35 % of the free host memory per node
            import psutil
            avail = psutil.virtual_memory().available
            per_stream = CHANNELS * (n_samples + n_hops * (T_BINS + N.N_METERS)) * 4
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            while es > 16 and es * per_stream * local_world > 0.35 * avail:
                es //= 2
        except Exception:
            pass
        hx = hcomb = hmet = None
        while True:                                   # biggest pinned host batch this box can give us
            ech = es * CHANNELS
            try:
                hx = torch.empty((ech, n_samples), dtype=torch.float32, pin_memory=True)
                hcomb = torch.empty((ech, n_hops, T_BINS), dtype=torch.float32, pin_memory=True)
                hmet = torch.empty((ech, n_hops, N.N_METERS), dtype=torch.float32, pin_memory=True)
                break
            except RuntimeError:
                hx = hcomb = hmet = None
                if es <= 16:
                    raise
                es //= 2
        calls = (n_streams + es - 1) // es
        hx.copy_(x[:ech])

        def e2e_step():
            for _ in range(calls):
                rc = N.lib().omega4_analyze(plan.handle, None, N.MEM_HOST, hx.data_ptr(), n_samples, ech, n_hops, 0,
                                            hcomb.data_ptr(), None, hmet.data_ptr(), None, None, None, 0)
                N.check(rc, "omega4_analyze(host)")
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        et = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev)
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        h2d = calls * hx.numel() * 4
        d2h = calls * (hcomb.numel() + hmet.numel()) * 4
        e2e = {"value": calls * es * args.seconds * world / float(et[0]), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * float(et[0]),
               "host_buffers": f"{calls} calls x {es} streams x {args.seconds} s per step from the same pinned buffers",
               "parity_vs_resident": bool(torch.equal(hmet, met[:ech].cpu()))}
        del hx, hcomb, hmet

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        ch_hops = n_ch * n_hops
        dom = max(ktimes, key=ktimes.get) if ktimes else None
        roof = None
        if dom:
            # algorithmic bytes of the dominant kernel per launch: every input sample once + what it writes
            out_bytes = {"multires_fft_8192": 6 * 4, "multires_fft_4096": 20 * 4, "multires_fft_2048": 102 * 4,
                         "multires_fft_1024": 384 * 4, "true_peak": 8, "kweight_lufs": 8, "meter_stats": 20,
                         "blockdft_gemm": 128 * 4, "blockdft_tc_gemm": 512 * 4, "blockdft_asm_8192": 6 * 4,
                         "blockdft_asm_4096": 20 * 4}.get(dom, 0)
            in_bytes = {"meter_stats": 16, "blockdft_asm_8192": 100 * 4, "blockdft_asm_4096": 400 * 4}.get(dom, HOP * 4)
            alg = ch_hops * (in_bytes + out_bytes)
            ach = alg / (ktimes[dom] / 1e3) / 1e9
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                    tj = json.load(f)
                if tj.get("kernel") == dom and tj.get("channel_hops") == ch_hops:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                pass
            # nominal arithmetic of the kernel per channel-hop (5 N log2 N per complex FFT, FMA = 2): the compute-side
            # view of the same kernel, next to the HBM view the contract asks for
            kflop = {"true_peak": 2 * 5 * 2048 * 11 + 3 * 2048 * 6, "kweight_lufs": 4 * 2080 * 14,
                     "multires_fft_2048": 5 * 1024 * 10 + 1024 * 12 + 2048, "multires_fft_1024": 5 * 512 * 9 + 512 * 12 + 1024,
                     "multires_fft_4096": 5 * 2048 * 11 + 2048 * 12 + 4096, "multires_fft_8192": 5 * 4096 * 12 + 4096 * 12 + 8192}.get(dom)
            fp32 = None
            if kflop:
                tf = ch_hops * kflop / (ktimes[dom] / 1e3) / 1e12
                fp32 = {"flop_per_channel_hop": kflop, "achieved_tflops": tf, "peak_tflops_measured": FP32_PEAK_TFLOPS,
                        "frac": tf / FP32_PEAK_TFLOPS,
                        "note": "FFT butterflies are additions, not FMAs: at 100 % issue this instruction mix reaches about half the FMA peak; "
                                "ncu smsp__issue_active of this kernel is 69 % (profiles/)"}
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": traffic, "peak_source": peak_kind, "kernel_ms": ktimes[dom], "fp32": fp32,
                    "kernel_share_of_step": ktimes[dom] / sum(ktimes.values()),
                    "note": "fp32-issue bound by arithmetic (about 220 FLOP per compulsory byte): low HBM fraction is expected; "
                            "kernel_share_of_step is of the summed kernel times (the statistics kernel overlaps the FFT kernels)"}
        pipe_gbs = ch_hops * B_ALG_PER_HOP / (ms_per_step / 1e3) / 1e9
        line = {
            "metric": METRIC if args.workload == "config2" else METRIC.replace("48 kHz stereo", "96 kHz 8-channel"),
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (FFT, combine, true peak, K-weighting in a float32 delta-form state; 3xTF32 tensor-core GEMM for the few-bin resolutions) + f64 (meter statistics)",
            "data": "synthetic",
            "config": {"workload": f"{wl_name}: {n_streams} streams x {CHANNELS} ch x {args.seconds} s @{SR // 1000} kHz per GPU, "
                                   f"resolutions {'/'.join(str(c[1]) for c in configs)}, hop 512, 512 target bins (fused output mode B), "
                                   "LUFS M/S/I/LRA + 4x true peak per hop",
                       "streams_per_gpu": n_streams, "seconds": args.seconds, "channels": CHANNELS,
                       "channel_hops_per_step_per_gpu": ch_hops, "parallelism": f"streams sharded x{world}, no data-path collective",
                       "l2": f"inputs {x.numel()*4/1e9:.1f} GB + outputs {comb.numel()*4/1e9:.1f} GB per step >> 126 MB L2 (no flush needed)"},
            "roofline": roof,
            "pipeline_hbm": {"algorithmic_bytes_per_channel_hop": B_ALG_PER_HOP, "achieved_gbs": pipe_gbs, "frac": pipe_gbs / hbm_peak},
            "roofline_fp32": {"flop_per_channel_hop": FLOP_PER_HOP, "achieved_tflops": ch_hops * FLOP_PER_HOP / (ms_per_step / 1e3) / 1e12,
                              "peak_tflops_measured": FP32_PEAK_TFLOPS,
                              "frac": ch_hops * FLOP_PER_HOP / (ms_per_step / 1e3) / 1e12 / FP32_PEAK_TFLOPS},
            "kernel_ms": ktimes, "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "next_rows": next_rows,
            "final_rows_checksum": checksum,
        }
        if world == 1 and not args.no_cpu:
            try:
                res, cores, n = cpu_baseline(args.cpu_seconds)
                line["cpu_baseline"] = {"value": res[0]["value"], "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"{n} stereo streams x {args.cpu_seconds:g} s, {cores} processes, OMP_NUM_THREADS=1, "
                                                  f"{res[0]['wall_s']:.1f} s wall",
                                        "cpu_s_per_stream_second_per_core": res[0]["cpu_s_per_stream_second"]}
            except Exception as e:       # the GPU numbers stand on their own
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
