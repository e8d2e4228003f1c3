#!/bin/bash
# timing experiments (wrong results) on the tensor-core hop-block GEMM: which part bounds it?
for X in "" "-DTC_EXPERIMENT_NO_LDG" "-DTC_EXPERIMENT_KS=1" "-DTC_EXPERIMENT_NO_EPI" "-DTC_EXPERIMENT_NO_LDG -DTC_EXPERIMENT_NO_EPI -DTC_EXPERIMENT_B_DIV=8"; do
  OMEGA4_NVCC_EXTRA="$X" python audio-analyzer-omega_b200/build.py --force > /dev/null
  timeout 300 python bench.py --steps 3 --warmup 3 --streams 128 --seconds 20 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$X]', 'gemm', d['kernel_ms']['blockdft_tc_gemm'])"
done
