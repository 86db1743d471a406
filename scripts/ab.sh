# same-box A/B of library builds / env settings
run() {
env $1 CASA_LIB_PATH=$PWD/casapose_b200/csrc/$2 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'score_ms', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],3))"
}
for i in 1 2; do
for lib in "$@"; do run X=0 $lib; done
done
