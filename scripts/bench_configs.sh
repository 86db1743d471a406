# Bench lines of every BASELINE.json config on one GPU (config 4 at N = 1 is the strong-scaling base line).
# usage (on a B200): bash scripts/bench_configs.sh [out dir, default gpurun_out]
O=${1:-gpurun_out}; mkdir -p $O
run() { name=$1; shift; timeout 600 python bench.py "$@" > $O/r02_bench_$name.json 2> $O/r02_bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/r02_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"]; e=d.get("e2e") or {}
    print("%-14s value %9.0f fr/s  %.3f ms/step  k_score %.3f ms frac %.3f  frame_frac %.3f  e2e %s" % ("$name", d["value"], d["ms_per_step"], r["launch_ms"] or 0, r["frac"] or 0, r["frame_frac"] or 0, ("%.0f" % e["value"]) if e else "-"))
except Exception as ex:
    print("$name FAILED", ex); print(open("$O/r02_bench_$name.err").read()[-800:])
PY
}
run cfg2 --config 2 --steps 200 --warmup 5 --no-cpu-baseline
run cfg2_multi --config 2 --variant multi --steps 100 --warmup 5 --no-cpu-baseline
run cfg3 --config 3 --steps 100 --warmup 5 --no-cpu-baseline
run cfg3_multi --config 3 --variant multi --steps 50 --warmup 5 --no-cpu-baseline
run cfg4_n1 --config 4 --steps 20 --warmup 3 --no-cpu-baseline
for hn in 128 256 512 1024 2048; do run cfg5_hn$hn --config 5 --hn $hn --steps 100 --warmup 5 --no-cpu-baseline; done
