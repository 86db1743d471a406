# same-box A/B of library builds: bash scripts/ab.sh <so under casapose_b200/csrc/> ...   (2 passes, 60 steps each)
run() {
CASA_LIB_PATH=$PWD/casapose_b200/csrc/$1 python bench.py --steps 60 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'score_ms', round(d['roofline']['launch_ms'],4), 'frac', round(d['roofline']['frac'],4))"
}
for i in 1 2; do
for lib in "$@"; do run $lib; done
done
