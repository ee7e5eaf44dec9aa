set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
: > gpurun_out/exp8.log
B="timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { echo "## $*" >> gpurun_out/exp8.log; env "$@" 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/exp8.log; }
run X=1 $B --workload config3-dpw-gridworld
run MCTSB200_LANE_ORDER=0 $B --workload config3-dpw-gridworld
run X=1 $B --workload config3-dpw-gridworld --roots 65536
run MCTSB200_FAST=0 $B --workload config3-dpw-gridworld
cat gpurun_out/exp8.log
