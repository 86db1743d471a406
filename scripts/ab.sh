for i in 1 2; do
for lib in libprev.so libcasapose_b200.so; do
CASA_LIB_PATH=$PWD/casapose_b200/csrc/$lib python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'score_ms', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],3))"
done; done
