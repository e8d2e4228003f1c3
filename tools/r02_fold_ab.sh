# same-box A/B at the full configuration: hop-block scales folded into the K-weighting kernel (default) vs the separate pass
for i in 1 2; do
for m in fold separate; do
  if [ $m = separate ]; then export OMEGA4_NO_SCALE_FOLD=1; else unset OMEGA4_NO_SCALE_FOLD; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms']
print('$m', 'step %.2f' % d['ms_per_step'], 'chk %.6f' % d['final_rows_checksum'], ' '.join('%s=%.2f' % (n, k[n]) for n in k if n in ('kweight_lufs','true_peak','multires_fft_2048','blockdft_row_scale','blockdft_tc_gemm')), d['clocks'])"
done; done
