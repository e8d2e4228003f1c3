mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_stress.py 2>&1 | tail -3 > gpurun_out/r02f_tests.txt
python tests/tools/random_stress.py 3 tc config2 >> gpurun_out/r02f_tests.txt 2>&1
python bench.py --steps 5 --warmup 3 --streams 128 --seconds 20 --no-cpu --no-e2e > gpurun_out/r02f_small.json 2>&1
python bench.py --steps 5 --no-cpu --no-e2e > gpurun_out/r02f_full.json 2>&1
cat gpurun_out/r02f_tests.txt
python - <<'PY'
import json
for f in ("gpurun_out/r02f_small.json","gpurun_out/r02f_full.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["kernel_ms"].items()})
PY
