# all-half true-peak kernel: smoke, A/B against the mixed kernel on one box, GPU tests, random stress (true-peak bar 0.05 dBTP)
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 || { echo "SMOKE FAILED/TIMED OUT"; exit 1; }
for i in 1 2; do
for m in half mixed; do
  if [ $m = mixed ]; then export OMEGA4_TP_MIXED=1; else unset OMEGA4_TP_MIXED; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --streams 256 --seconds 30 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms']
print('$m', 'step %.2f' % d['ms_per_step'], ' '.join('%s=%.2f' % (n, k[n]) for n in ('kweight_lufs','true_peak','multires_fft_2048')))"
done; done 2>&1 | tee gpurun_out/r02l_ab.txt
unset OMEGA4_TP_MIXED
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^FAILED|passed|failed|Error" | head -20 | tee gpurun_out/r02l_tests.txt
for s in 0 1 2 3 4 5; do echo "seed $s tc config2: $(timeout 120 python tests/tools/random_stress.py $s tc config2 2>&1 | tail -1 | cut -c1-200)"; done | tee gpurun_out/r02l_stress.txt
for s in 0 1; do echo "seed $s tc config5: $(timeout 200 python tests/tools/random_stress.py $s tc config5 2>&1 | tail -1 | cut -c1-200)"; done | tee -a gpurun_out/r02l_stress.txt
