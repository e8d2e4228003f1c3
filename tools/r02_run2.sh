mkdir -p gpurun_out
python -m pytest tests -m gpu -q --deselect tests/test_gpu_stress.py 2>&1 | tail -40 > gpurun_out/r02b_tests.txt
python tools/stream_latency.py > gpurun_out/r02b_stream_latency.txt 2>&1
python tools/pcie_bw.py > gpurun_out/r02b_pcie.txt 2>&1
python bench.py --steps 5 --e2e-variants bars,meters,f32 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
python bench.py --workload config3 --steps 1 --no-cpu --e2e-variants "" > gpurun_out/r02b_bench_c3.json 2> gpurun_out/r02b_bench_c3.err
tail -5 gpurun_out/r02b_tests.txt; tail -3 gpurun_out/r02b_bench.err gpurun_out/r02b_bench_c3.err
