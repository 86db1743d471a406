# Compiles and runs the stand-alone tcgen05 probes under short timeouts.  Not part of the library, the tests or the bench.
# usage (on a B200):  bash scripts/tc_probe.sh            descriptor / layout probe (one MMA, host check; both LBO/SBO orders)
#                     bash scripts/tc_probe.sh epilogue   epilogue-rate probe (run the first one before it)
mkdir -p gpurun_out
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
if [ "$1" = "epilogue" ]; then
  nvcc $FLAGS -o /tmp/tc_epilogue_probe casapose_b200/csrc/experimental/tc_epilogue_probe.cu || exit 1
  timeout 30 /tmp/tc_epilogue_probe ${2:-4000}
  echo "exit $?"
  nvcc $FLAGS -DCNT_PRMT -o /tmp/tc_epilogue_probe_prmt casapose_b200/csrc/experimental/tc_epilogue_probe.cu || exit 1
  timeout 30 /tmp/tc_epilogue_probe_prmt ${2:-4000}
  echo "exit $?"
else
  nvcc $FLAGS -o /tmp/tc_probe casapose_b200/csrc/experimental/tc_probe.cu || exit 1
  timeout 20 /tmp/tc_probe
  echo "exit $?"
  timeout 20 /tmp/tc_probe swap
  echo "exit $?"
fi
