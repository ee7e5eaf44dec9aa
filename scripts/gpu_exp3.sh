set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
: > gpurun_out/exp3.log
B="timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline"
run() { echo "## $*" >> gpurun_out/exp3.log; env "$@" 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/exp3.log; }
run X=1 $B
run MCTSB200_LANE_ORDER=1 $B
run X=1 $B --workload config2-uct-gridworld-stochastic
run MCTSB200_LANE_ORDER=1 $B --workload config2-uct-gridworld-stochastic
cat gpurun_out/exp3.log
