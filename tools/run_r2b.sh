for v in 1 2 3 4 5 6; do
  python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('run$v', 'ms/step', round(d['ms_per_step'],3), 'host', round(d['host_issue_ms_per_step'],3), d['clocks']['samples'], {k:(round(v['ms_per_step'],3), round(v['host_issue_ms_per_step'],3)) for k,v in d['config']['other_routes'].items()})"
done
