mkdir -p gpurun_out
(lscpu | head -30; echo; cat /sys/devices/system/node/node*/cpulist 2>/dev/null; echo; nvidia-smi topo -m; echo; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q 0x0302 $d/class 2>/dev/null; then echo $d $(cat $d/numa_node); fi; done; free -g; which numactl) > gpurun_out/r02_topo.txt 2>&1
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_stress.py 2>&1 | tail -40 > gpurun_out/r02a_tests.txt
python -m pytest tests/test_gpu_stress.py -m gpu -q 2>&1 | tail -15 >> gpurun_out/r02a_tests.txt
for s in 0 1 2; do for m in tc fft; do echo "seed $s $m config5: $(python tests/tools/random_stress.py $s $m config5 2>&1 | tail -1)"; done; done > gpurun_out/r02a_stress96.txt 2>&1
for s in 0 1 2 3; do echo "seed $s tc config2: $(python tests/tools/random_stress.py $s tc config2 2>&1 | tail -1)"; done >> gpurun_out/r02a_stress96.txt 2>&1
tail -5 gpurun_out/r02a_tests.txt
