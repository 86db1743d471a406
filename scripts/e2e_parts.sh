for parts in 1 2 4 8; do
CASA_HOST_PARTS=$parts timeout 200 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('parts $parts', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2))"
done
