#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/spread.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for w in "config3-dpw-gridworld" "config2-uct-gridworld-stochastic --roots 16384" "config2-uct-gridworld-default --roots 8192" "config3-dpw-gridworld --roots 4096"; do
  for sp in 1 2 4 8; do
    echo -n "SPREAD=$sp $w: " >> gpurun_out/spread.log
    MCTSB200_SPREAD=$sp timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $w 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/spread.log
  done
done
cat gpurun_out/spread.log
