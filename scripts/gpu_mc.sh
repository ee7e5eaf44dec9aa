#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/mc.log
B="timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload config4-dpw-mountaincar"
echo -n "default: " >> gpurun_out/mc.log; $B 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/mc.log
for sp in 32 16 8; do
  echo -n "lanes1 SPREAD=$sp: " >> gpurun_out/mc.log; MCTSB200_SPREAD=$sp $B --lanes 1 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/mc.log
done
echo -n "lanes1 nospread: " >> gpurun_out/mc.log; MCTSB200_SPREAD=1 $B --lanes 1 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/mc.log
cat gpurun_out/mc.log
