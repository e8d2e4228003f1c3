# persistent GEMM: quick liveness check, timeline, A/B, then parity (each step under its own timeout)
V=${1:-p16}
L=$PWD/build/variants/$V.so
OMEGA4_CUDA_LIB=$L timeout 120 python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 || { echo "SMOKE FAILED/TIMED OUT"; exit 1; }
OMEGA4_CUDA_LIB=$PWD/build/variants/${V}_tl.so timeout 200 python bench.py --steps 1 --warmup 1 --streams 256 --seconds 30 --no-cpu --no-e2e 2>&1 | grep tc_timeline | tail -12
tools/ab_variants.sh gpurun_out/ab_$V.txt f16d $V ptf32 $V | sed -E "s/kweight_lufs.*multires_fft_1024=[0-9.]+ //"
OMEGA4_CUDA_LIB=$L timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for s in 0 1 2; do echo "seed $s tc config2: $(OMEGA4_CUDA_LIB=$L timeout 120 python tests/tools/random_stress.py $s tc config2 2>&1 | tail -1 | cut -c1-150)"; done
for s in 0; do echo "seed $s tc config5: $(OMEGA4_CUDA_LIB=$L timeout 200 python tests/tools/random_stress.py $s tc config5 2>&1 | tail -1 | cut -c1-150)"; done
