# Round-end evidence on one B200: tests, bench lines, ncu launch lists.  usage: bash scripts/round_end.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 2500 gpurun_out/bench_r01.json
timeout 200 python bench.py --workload ls > gpurun_out/bench_ls_r01.json 2>/dev/null; tail -c 1200 gpurun_out/bench_ls_r01.json
timeout 200 python scripts/gpu_next_rows.py > gpurun_out/next_rows_r01.txt 2>&1; cat gpurun_out/next_rows_r01.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/r01_launches.csv python scripts/gpu_step.py 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_next_rows_launches.csv python scripts/gpu_next_rows.py 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_mask_bits -s 1 -c 1 --page raw --csv --log-file gpurun_out/r01_k_mask_bits_full.csv python scripts/gpu_step.py 3 > /dev/null 2>&1
echo done
