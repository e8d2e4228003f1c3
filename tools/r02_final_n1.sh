#!/bin/bash
# one-GPU records of the final build: config[3] weak (10-minute streams), config 5, config[3] strong point at N = 1, stream latency
TAG=${1:-r02l}
mkdir -p gpurun_out
python bench.py --workload config3 --steps 2 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_config3.json 2> gpurun_out/${TAG}_bench_config3.err
python bench.py --workload config5 --streams 64 --seconds 30 --steps 5 --e2e-variants "" > gpurun_out/${TAG}_bench_c5_n1.json 2> gpurun_out/${TAG}_bench_c5_n1.err
python bench.py --workload config3 --total-streams 8192 --steps 1 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_c3strong_n1.json 2> gpurun_out/${TAG}_bench_c3strong_n1.err
python tools/stream_latency.py > gpurun_out/${TAG}_stream_latency.txt 2>&1
for f in config3 c5_n1 c3strong_n1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${f}.json").read().strip().splitlines()[-1])
    print("$f", "value %.1f ms %.2f e2e %s oracle %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), d.get("oracle_check")))
    print({k: round(v, 2) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print("$f FAILED", e)
PY
done
cat gpurun_out/${TAG}_stream_latency.txt | tail -5
