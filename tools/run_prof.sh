#!/bin/bash
# usage: tools/run_prof.sh TAG   -- full bench, small plain bench, ncu launch list, ncu --set full of one step,
#                                   ncu --set full of the dominant kernel at the full default configuration
TAG=${1:-r01x}
python bench.py > gpurun_out/${TAG}_full.log 2> gpurun_out/${TAG}_full.err
python bench.py --steps 3 --warmup 3 --streams 128 --seconds 20 --no-cpu --no-e2e > gpurun_out/${TAG}_small_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --streams 128 --seconds 20 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -s 25 -c 8 -o gpurun_out/prof_${TAG} -f python bench.py --steps 1 --warmup 3 --streams 128 --seconds 20 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:truepeak -s 3 -c 1 -o gpurun_out/prof_${TAG}_tp_full -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_tp_full.log 2>&1
python tools/parity_report.py > gpurun_out/${TAG}_parity.txt 2>&1
python tools/tone_stress.py > gpurun_out/${TAG}_tone_stress.txt 2>&1
for s in 0 1 2 3 4 5; do for m in tc fft tcfd fp32; do echo "seed $s $m: $(python tests/tools/random_stress.py $s $m 2>&1 | tail -1)"; done; done > gpurun_out/${TAG}_random_stress.txt 2>&1
tail -c 400 gpurun_out/${TAG}_full.log
