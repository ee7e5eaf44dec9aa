set -x
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
for L in 1 2 4 8 32; do timeout 600 python bench.py --steps 2 --warmup 1 --lanes $L --no-cpu-baseline; done 2>&1 | tee gpurun_out/lanes.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tee gpurun_out/bench.log
