mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -m gpu -q -k "adversarial" 2>&1 | grep -v "^  " | grep "Error\|^E  \|assert\|passed\|failed" | head -20 > gpurun_out/r02e_tests.txt
(nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|^CPU(s)"; free -g | head -2) > gpurun_out/r02e_topo2.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_bw.py > gpurun_out/r02e_pcie_n2.json 2> gpurun_out/r02e_pcie_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/pcie_bw.py --no-numa > gpurun_out/r02e_pcie_n2_nonuma.json 2>> gpurun_out/r02e_pcie_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --e2e-variants "" > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err
cat gpurun_out/r02e_tests.txt; tail -n 3 gpurun_out/r02e_bench_n2.err
