#!/bin/bash
# developer sweep: rebuild with different KW32_MIN_BLOCKS on the GPU box and time the K-weighting kernel
for MB in 3 4 5 6; do
  OMEGA4_NVCC_EXTRA="-DKW32_MIN_BLOCKS=$MB" python audio-analyzer-omega_b200/build.py --force -v 2>&1 | grep -A2 "kweight32" | grep -E "spill|Used"
  python bench.py --steps 3 --warmup 3 --streams 128 --seconds 20 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MIN_BLOCKS $MB', 'kweight', d['kernel_ms']['kweight_lufs'], 'step', d['ms_per_step'])"
done
