#!/bin/bash
# usage: tools/r02_multi.sh N TAG  -- multi-GPU evidence: topology, concurrent copy probe, config2 (weak), config5, config3 strong scaling
N=${1:-8}; TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|^CPU(s)\|model name"; free -g | head -2; for d in /sys/bus/pci/devices/*; do if grep -q 0x0302 $d/class 2>/dev/null; then echo $d numa_node $(cat $d/numa_node); fi; done) > gpurun_out/${TAG}_topo_n$N.txt 2>&1
if [ "$N" = "1" ]; then TR="python"; fi
$TR ${N:+$( [ "$N" != "1" ] && echo --master-port 29521 )} tools/pcie_bw.py > gpurun_out/${TAG}_pcie_n$N.json 2> gpurun_out/${TAG}_pcie_n$N.err
$TR $( [ "$N" != "1" ] && echo --master-port 29522 ) bench.py --gpus $N --steps 3 --e2e-variants "" --no-cpu > gpurun_out/${TAG}_bench_c2_n$N.json 2> gpurun_out/${TAG}_bench_c2_n$N.err
$TR $( [ "$N" != "1" ] && echo --master-port 29523 ) bench.py --gpus $N --workload config5 --streams 64 --seconds 30 --steps 3 --e2e-variants "" > gpurun_out/${TAG}_bench_c5_n$N.json 2> gpurun_out/${TAG}_bench_c5_n$N.err
$TR $( [ "$N" != "1" ] && echo --master-port 29524 ) bench.py --gpus $N --workload config3 --total-streams 8192 --steps 1 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_c3strong_n$N.json 2> gpurun_out/${TAG}_bench_c3strong_n$N.err
for f in gpurun_out/${TAG}_bench_c2_n$N.json gpurun_out/${TAG}_bench_c5_n$N.json gpurun_out/${TAG}_bench_c3strong_n$N.json; do tail -c 200 $f; echo; done
tail -n 2 gpurun_out/${TAG}_bench_c3strong_n$N.err
