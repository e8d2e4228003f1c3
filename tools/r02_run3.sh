mkdir -p gpurun_out
python -m pytest tests -m gpu -q --deselect tests/test_gpu_stress.py 2>&1 | grep -v "^  " | grep "Error\|^E  \|assert\|passed\|failed" | head -60 > gpurun_out/r02d_tests.txt
cat gpurun_out/r02d_tests.txt | tail -20
