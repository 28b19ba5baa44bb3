echo "== NW=12 TB=24"; timeout 120 python tools/quick_bench.py 2>&1 | grep -E "backward"
echo "== NW=8 TB=16 NS=5"; ASG_GRAD_TB=16 ASG_GRAD_NS=5 timeout 120 python tools/quick_bench.py 2>&1 | grep -E "backward"
echo "== NW=8 TB=16 NS=3"; ASG_GRAD_TB=16 ASG_GRAD_NS=3 timeout 120 python tools/quick_bench.py 2>&1 | grep -E "backward"
