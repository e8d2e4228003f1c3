#!/bin/bash
# usage: tools/r02_final_multi.sh N TAG  -- device-resident multi-GPU records of the final build: config[2] weak, config 5, config[3] strong
N=${1:-8}; TAG=${2:-r02l}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29522 bench.py --gpus $N --steps 5 --no-e2e --no-cpu > gpurun_out/${TAG}_bench_c2_n$N.json 2> gpurun_out/${TAG}_bench_c2_n$N.err
$TR --master-port 29523 bench.py --gpus $N --workload config5 --streams 64 --seconds 30 --steps 5 --no-e2e --no-cpu > gpurun_out/${TAG}_bench_c5_n$N.json 2> gpurun_out/${TAG}_bench_c5_n$N.err
$TR --master-port 29524 bench.py --gpus $N --workload config3 --total-streams 8192 --steps 1 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_c3strong_n$N.json 2> gpurun_out/${TAG}_bench_c3strong_n$N.err
for f in c2 c5 c3strong; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${f}_n$N.json").read().strip().splitlines()[-1])
    print("$f", "value %.1f ms %.2f n_gpus %d parity %s oracle %s" % (d["value"], d["ms_per_step"], d["n_gpus"], d.get("multi_gpu_parity"), d.get("oracle_check")))
except Exception as e:
    print("$f FAILED", e)
PY
done
tail -n 3 gpurun_out/${TAG}_bench_c3strong_n$N.err
