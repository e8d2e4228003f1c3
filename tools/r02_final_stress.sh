# random adversarial stress of the final build (default path), both configurations; tone-leak probe
mkdir -p gpurun_out
{ for s in 0 1 2 3 4 5 6 7; do echo "seed $s tc config2: $(timeout 150 python tests/tools/random_stress.py $s tc config2 2>&1 | tail -1)"; done
  for s in 0 1 2; do echo "seed $s tc config5: $(timeout 250 python tests/tools/random_stress.py $s tc config5 2>&1 | tail -1)"; done
  for s in 0 1; do echo "seed $s fft config2: $(timeout 150 python tests/tools/random_stress.py $s fft config2 2>&1 | tail -1)"; done; } > gpurun_out/r02o_random_stress.txt 2>&1
cut -c1-170 gpurun_out/r02o_random_stress.txt
