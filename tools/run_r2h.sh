for e in "X=1" "ASG_GRAD_TB=16" "ASG_GRAD_TB=8"; do
echo "== $e"; env $e timeout 200 python tools/dbg_fac_state.py 1000 16 40 200 varlen 2>&1 | grep -E "gx rel err|gx diff" | cut -c1-200
done
