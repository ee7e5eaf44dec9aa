#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/order.log
for w in config2-uct-gridworld-default config2-uct-gridworld-stochastic config3-dpw-gridworld; do
  for o in 0 1; do
    echo -n "LANE_ORDER=$o $w: " >> gpurun_out/order.log
    MCTSB200_LANE_ORDER=$o timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $w 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/order.log
  done
done
cat gpurun_out/order.log
