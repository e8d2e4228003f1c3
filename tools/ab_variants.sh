#!/bin/bash
# A/B runs of pre-built library variants (tools/build_variant.sh) on one box: per-kernel ms of a reduced config[2]
# usage: tools/ab_variants.sh OUTFILE name1 name2 ...
out=$1; shift
: > $out
for v in "$@"; do
  OMEGA4_CUDA_LIB=$PWD/build/variants/$v.so timeout 300 python bench.py --steps 3 --warmup 3 --streams ${AB_STREAMS:-256} --seconds ${AB_SECONDS:-30} --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms']
print('$v', 'step %.2f' % d['ms_per_step'], 'chk %.6f' % d['final_rows_checksum'], ' '.join('%s=%.2f' % (n, k[n]) for n in k))" >> $out 2>&1 || echo "$v FAILED" >> $out
done
cat $out
