# usage: tools/r02_check.sh TAG  -- GPU tests + default bench (BASELINE config[2], e2e + CPU leg) on one box
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^FAILED|^ERROR|passed|failed" | head -20 | tee gpurun_out/${TAG}_tests.txt
timeout 900 python bench.py --steps 5 --e2e-variants bars,meters > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.1f ms %.2f e2e %.1f (%.1f ms) cpu %s frac %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"], d["roofline"]["frac"]))
print({k: round(v, 2) for k, v in d["kernel_ms"].items()})
PY
