# Instruction-mix benchmark of the k_score inner loop (design exploration, not part of the library).
# usage (on a B200):  bash scripts/loop_bench.sh [resident blocks per SM, default 3]
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
nvcc $FLAGS -o /tmp/loop_bench casapose_b200/csrc/experimental/loop_bench.cu 2>/dev/null || exit 1
timeout 60 /tmp/loop_bench ${1:-3}
