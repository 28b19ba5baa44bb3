set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2i_tests.log
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "host", round(d["host_issue_ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "lat", round(d["e2e"]["latency_ms_one_call"], 3))
print({k: round(v["avg_us"], 1) for k, v in d["kernels"].items()}, "roofline", round(d["roofline"]["frac"], 3))
print(d["config"]["other_routes"]); print(d["clocks"]); print(d.get("cpu_baseline", {}).get("value"), d.get("gpu_reference"))
PY
