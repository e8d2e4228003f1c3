#!/bin/bash
# usage: tools/run_prof.sh TAG   -- small plain bench, ncu launch list, ncu --set full of one step (all kernels, with source),
#                                   ncu --set full of the dominant kernel at the full default configuration
TAG=${1:-r02x}
mkdir -p gpurun_out
SMALL="--steps 1 --warmup 3 --streams 128 --seconds 20 --no-cpu --no-e2e"
KREG='regex:kweight32|truepeak|stats_kernel|multires_local|hopblock|blockdft'
python bench.py --steps 3 --warmup 3 --streams 128 --seconds 20 --no-cpu --no-e2e > gpurun_out/${TAG}_small_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --streams 128 --seconds 20 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_list.log 2>&1
# one step's kernels: skip the three warm-up steps' launches of the same kernels (9 per step)
ncu --set full --clock-control none --import-source on -k "$KREG" -s 27 -c 9 -o gpurun_out/prof_${TAG} -f python bench.py $SMALL > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:truepeak -s 3 -c 1 -o gpurun_out/prof_${TAG}_tp_full -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_tp_full.log 2>&1
tail -c 300 gpurun_out/${TAG}_small_plain.log
