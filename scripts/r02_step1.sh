# round-2 GPU step: parity tests, bench, ncu of k_score (source regions)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_tests_a.log; cat gpurun_out/r02_tests_a.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -c 1800 gpurun_out/r02_bench_a.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_score -s 1 -c 1 -o gpurun_out/r02_score_a python scripts/gpu_step.py 2 > /dev/null 2>&1
ls -la gpurun_out/r02_score_a.ncu-rep
