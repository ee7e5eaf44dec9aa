set -x
: > gpurun_out/exp2.log
B="timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline"
run() { echo "## $*" >> gpurun_out/exp2.log; env "$@" 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/exp2.log; }
run MCTSB200_LANE_ORDER=1 $B
run MCTSB200_LANE_ORDER=1 $B --workload config2-uct-gridworld-stochastic
cat gpurun_out/exp2.log
export MCTSB200_LANE_ORDER=1
bash scripts/gpu_phase_timing.sh " "
