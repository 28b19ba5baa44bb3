import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d["kernels"]
print(sys.argv[1], round(d["value"]), "us/step", round(d["ms_per_step"] * 1e3, 1), "host", round(d.get("host_issue_ms_per_step", 0) * 1e3, 1),
      "e2e", round(d["e2e"]["value"]), {n: round(v["avg_us"], 1) for n, v in k.items()})
