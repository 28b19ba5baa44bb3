timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "large or cfg4" 2>&1 | tail -4
timeout 200 python tools/quick_bench.py 500 128 1024 100 2>&1 | tail -8
timeout 500 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_cfg4.json 2> gpurun_out/r2j_cfg4.err; tail -3 gpurun_out/r2j_cfg4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2j_cfg4.json").read().strip().splitlines()[-1])
print("cfg4 value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), {k: round(v["avg_us"], 1) for k, v in d["kernels"].items()}, d.get("roofline", {}).get("frac"))
PY
