# A/B of the half-operand GEMM against the 3xTF32 one + parity of the half path (tests, random stress, tone leak)
L=$PWD/build/variants/${1:-f16}.so
tools/ab_variants.sh gpurun_out/ab_f16.txt tf32 ${1:-f16} tf32 ${1:-f16}
OMEGA4_CUDA_LIB=$L timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/f16_tests.txt
for s in 0 1 2 3 4 5; do echo "seed $s tc config2: $(OMEGA4_CUDA_LIB=$L timeout 120 python tests/tools/random_stress.py $s tc config2 2>&1 | tail -1)"; done > gpurun_out/f16_stress.txt 2>&1
for s in 0 1; do echo "seed $s tc config5: $(OMEGA4_CUDA_LIB=$L timeout 200 python tests/tools/random_stress.py $s tc config5 2>&1 | tail -1)"; done >> gpurun_out/f16_stress.txt 2>&1
OMEGA4_CUDA_LIB=$L timeout 200 python tests/tools/tone_leak_probe.py > gpurun_out/f16_tone_leak.txt 2>&1
cat gpurun_out/f16_tests.txt gpurun_out/f16_stress.txt; tail -20 gpurun_out/f16_tone_leak.txt
