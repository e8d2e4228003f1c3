#!/bin/bash
# usage: tools/r02_strong.sh N  -- config3 strong-scaling point (8192 streams x 10 min in total) on N GPUs
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --workload config3 --total-streams 8192 --steps 1 --no-cpu --no-e2e > gpurun_out/r02_bench_c3strong_n1.json 2> gpurun_out/r02_bench_c3strong_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --workload config3 --total-streams 8192 --steps 1 --no-cpu --no-e2e > gpurun_out/r02_bench_c3strong_n$N.json 2> gpurun_out/r02_bench_c3strong_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/pcie_bw.py > gpurun_out/r02_pcie_n$N.json 2> gpurun_out/r02_pcie_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --e2e-variants "" --no-cpu > gpurun_out/r02_bench_c2_n$N.json 2> gpurun_out/r02_bench_c2_n$N.err
fi
tail -c 300 gpurun_out/r02_bench_c3strong_n$N.json; tail -n 3 gpurun_out/r02_bench_c3strong_n$N.err
