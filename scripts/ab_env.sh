# same-box A/B of environment settings with the current library:  bash scripts/ab_env.sh "A=1" "B=2" ...
run() {
env $1 timeout 200 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'score_ms', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],3))"
}
for i in 1 2; do
for e in "$@"; do run $e; done
done
