set -x
: > gpurun_out/exp7.log
B="timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { echo "## $*" >> gpurun_out/exp7.log; env "$@" 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/exp7.log; }
V=$PWD/mcts.jl_b200/csrc/variants
run X=1 $B --workload config2-uct-gridworld-stochastic
run MCTSB200_LIB=$V/liboldorder.so $B --workload config2-uct-gridworld-stochastic
run MCTSB200_LIB=$V/liboldorder.so $B
run MCTSB200_LIB=$V/liboldorder.so MCTSB200_LANE_ORDER=0 $B --workload config2-uct-gridworld-stochastic
cat gpurun_out/exp7.log
