#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/episodes.log
for a in "--kind uct" "--kind uct --reuse" "--kind dpw --episodes 16384" "--kind dpw --episodes 16384 --reuse" "--kind dpw --episodes 65536"; do
  echo "## $a" >> gpurun_out/episodes.log
  timeout 900 python scripts/bench_episodes.py $a --cpu-episodes 64 --check >> gpurun_out/episodes.log 2>&1
done
cat gpurun_out/episodes.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
