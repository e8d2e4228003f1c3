mkdir -p gpurun_out
python -m pytest tests -m gpu -q --deselect tests/test_gpu_stress.py 2>&1 | grep -v "^  " | grep "Error\|^E  \|assert\|passed\|failed" | head -40 > gpurun_out/r02g_tests.txt
python tools/stream_latency.py > gpurun_out/r02g_stream_latency.txt 2>&1
for s in 0 1 2 3; do echo "seed $s tc config5: $(python tests/tools/random_stress.py $s tc config5 2>&1 | tail -1)"; done > gpurun_out/r02g_stress96.txt 2>&1
python bench.py --workload config5 --streams 64 --seconds 30 --steps 5 --e2e-variants "" > gpurun_out/r02g_bench_c5.json 2> gpurun_out/r02g_bench_c5.err
bash tools/r02_strong.sh 1
cat gpurun_out/r02g_tests.txt gpurun_out/r02g_stream_latency.txt gpurun_out/r02g_stress96.txt; tail -c 600 gpurun_out/r02g_bench_c5.json
