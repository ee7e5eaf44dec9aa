#!/bin/bash
# A/B of variant libraries (mcts.jl_b200/csrc/variants/lib<name>.so, "cur" = the in-tree build) on the bench workloads.
# usage: bash scripts/gpu_abv.sh "<name> <name> ..." "<workload> ..." [extra bench args]
mkdir -p gpurun_out; : > gpurun_out/abv.log
D=$PWD/mcts.jl_b200/csrc
for rep in 1 2; do
for w in $2; do
  for v in $1; do
    if [ "$v" = cur ]; then L=$D/libmctsb200.so; else L=$D/variants/lib$v.so; fi
    echo -n "$v $w: " >> gpurun_out/abv.log
    MCTSB200_LIB=$L timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $w $3 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/abv.log
  done
done
done
cat gpurun_out/abv.log
