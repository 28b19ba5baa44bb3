# developer A/B on one box: tools/ab.sh reps pkg [pkg ...]   (prefix a pkg with VAR=val: to set an env var)
reps=$1; shift
for i in $(seq 1 $reps); do for spec in "$@"; do
envs=""; p="$spec"; case "$spec" in *:*) envs="${spec%%:*}"; p="${spec#*:}";; esac
env $envs ASG_PKG=$PWD/$p ASG_BENCH_NCTX=8 python bench.py --steps 200 --warmup 10 --no-cpu-baseline | python tools/ab_fmt.py "$spec"
done; done
