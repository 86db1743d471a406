# Compiles and runs the stand-alone tcgen05 probe (casapose_b200/csrc/experimental/tc_probe.cu) under a short timeout.
# Not part of the library, the tests or the bench.  usage (on a B200):  bash scripts/tc_probe.sh
set -e
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o gpurun_out/tc_probe casapose_b200/csrc/experimental/tc_probe.cu
timeout 20 gpurun_out/tc_probe
