run() { echo "$@"; env "$@" timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e $LANES 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['launch_ms'],4))"; }
LANES="--lanes 3" run CASA_SCORE_BPS=3 CASA_NO_GRAPH=1
LANES="--lanes 3" run CASA_SCORE_BPS=4 CASA_NO_GRAPH=1
LANES="--lanes 4" run CASA_SCORE_BPS=3
LANES="--lanes 4" run CASA_SCORE_BPS=2
LANES="--lanes 4" run CASA_SCORE_BPS=4
LANES="--lanes 2" run CASA_SCORE_BPS=2
LANES="--lanes 2" run CASA_SCORE_BPS=3
LANES="--lanes 2" run CASA_SCORE_BPS=4
LANES="--lanes 3" run CASA_SCORE_BPS=1
