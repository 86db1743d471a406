# same-box A/B: calls in flight (lanes) x resident scoring blocks per SM (CASA_SCORE_BPS); results: profiles/r02_lanes_overlap.txt
run() { echo "$@"; env "$@" timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e $LANES 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['launch_ms'],4), round(d['roofline']['frac'],4), d['config'].get('ms_per_step_one_lane'))"; }
for bps in 4 3 2 1; do LANES="--lanes 3" run CASA_SCORE_BPS=$bps; done
for l in 2 4; do for bps in 4 3 2; do LANES="--lanes $l" run CASA_SCORE_BPS=$bps; done; done
