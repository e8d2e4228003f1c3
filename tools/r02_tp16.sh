V=${1:-tp16}
L=$PWD/build/variants/$V.so
OMEGA4_CUDA_LIB=$L timeout 120 python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 || { echo "SMOKE FAILED/TIMED OUT"; exit 1; }
tools/ab_variants.sh gpurun_out/ab_$V.txt cur $V cur $V | sed -E "s/multires_fft_2048.*//"
OMEGA4_CUDA_LIB=$L timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^FAILED|passed|failed|Error" | head -20
for s in 0 1 2 3 4 5; do echo "seed $s tc config2: $(OMEGA4_CUDA_LIB=$L timeout 120 python tests/tools/random_stress.py $s tc config2 2>&1 | tail -1 | cut -c1-150)"; done
for s in 0 1; do echo "seed $s tc config5: $(OMEGA4_CUDA_LIB=$L timeout 200 python tests/tools/random_stress.py $s tc config5 2>&1 | tail -1 | cut -c1-150)"; done
