#!/bin/bash
# developer sweep: producer groups (TC_PG) of the tensor-core hop-block GEMM
for PG in ${PGS:-1 2 4}; do
  OMEGA4_NVCC_EXTRA="-DTC_PG=$PG" python audio-analyzer-omega_b200/build.py --force > /dev/null
  timeout 300 python bench.py --steps 3 --warmup 3 --streams 128 --seconds 20 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TC_PG $PG', 'gemm', d['kernel_ms']['blockdft_tc_gemm'], 'step', d['ms_per_step'])"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "blockdft or baseline" 2>&1 | tail -2
