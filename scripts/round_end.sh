# Round-end evidence on one B200: tests, bench lines, ncu launch lists.  usage: bash scripts/round_end.sh   (outputs: gpurun_out/)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gpu_tests.log 2>&1; tail -3 gpurun_out/r02_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 3000 gpurun_out/r02_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
timeout 300 python bench.py --lanes 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_one_lane.json 2>/dev/null
timeout 300 python bench.py --workload ls > gpurun_out/r02_bench_ls.json 2>/dev/null; tail -c 1500 gpurun_out/r02_bench_ls.json
bash scripts/bench_configs.sh gpurun_out
# launch lists: direct launches (ncu cannot profile kernel nodes of graphs that may hold conditional nodes); cold = ncu's
# default cache control, warm = --cache-control none
CASA_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 10 --csv --log-file gpurun_out/r02_launches.csv python scripts/gpu_step.py 5 > /dev/null 2>&1
CASA_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 30 -c 10 --csv --log-file gpurun_out/r02_launches_warm.csv python scripts/gpu_step.py 5 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 18 -c 6 --csv --log-file gpurun_out/r02_launches_ls_warm.csv python scripts/gpu_ls_step.py 5 > /dev/null 2>&1
CASA_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none -k regex:k_score -s 1 -c 1 --page raw --csv --log-file gpurun_out/r02_k_score_full_raw.csv python scripts/gpu_step.py 3 > /dev/null 2>&1
echo done
