# Compiles and runs the stand-alone tcgen05 probes under short timeouts.  Not part of the library, the tests or the bench.
# usage (on a B200):  bash scripts/tc_probe.sh            descriptor / layout probe (one MMA, host check)
#                     bash scripts/tc_probe.sh epilogue   epilogue-rate probe (run the first one before it)
set -e
mkdir -p gpurun_out
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
if [ "$1" = "epilogue" ]; then
  nvcc $FLAGS -o gpurun_out/tc_epilogue_probe casapose_b200/csrc/experimental/tc_epilogue_probe.cu
  timeout 30 gpurun_out/tc_epilogue_probe ${2:-4000}
else
  nvcc $FLAGS -o gpurun_out/tc_probe casapose_b200/csrc/experimental/tc_probe.cu
  timeout 20 gpurun_out/tc_probe
fi
