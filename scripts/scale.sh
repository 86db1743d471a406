# Scaling run on one box: bash scripts/scale.sh "1 8" [extra bench args]   (inside gpurun --gpus 8)
mkdir -p gpurun_out
NS=${1:-"1 2 4 8"}; shift
for n in $NS; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 200 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_scale_n$n.json").read().strip().splitlines()[-1])
    print("N=$n value %.0f fr/s  %.4f ms/step  k_score %.4f ms  e2e %.0f fr/s (h2d %.1f GB/s)  verified %s" % (d["value"], d["ms_per_step"], d["roofline"]["launch_ms"], d["e2e"]["value"], d["e2e"]["h2d_gbs_measured"], d["config"]["parallelism"][-5:]))
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/r02_scale_n$n.err").read()[-1500:])
PY
done
