#!/bin/bash
bash scripts/gpu_quick.sh
for a in "--kind uct" "--kind dpw --episodes 16384"; do
  timeout 900 python scripts/bench_episodes.py $a --cpu-episodes 64 --check 2>&1 | tail -1 | cut -c1-420
done
