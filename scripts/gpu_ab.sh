# A/B bench lines for the tuning switches (appended to gpurun_out/ab.log), after the GPU parity tests
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
: > gpurun_out/ab.log
B="timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline"
run() { echo "## $*" >> gpurun_out/ab.log; env "$@" 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/ab.log; }
run X=1 $B
run MCTSB200_LANE_ORDER=0 $B
run MCTSB200_PATH_SMEM=0 $B
run MCTSB200_LANE_ORDER=0 MCTSB200_PATH_SMEM=0 $B
run X=1 $B --workload config2-uct-gridworld-stochastic
run MCTSB200_LANE_ORDER=0 $B --workload config2-uct-gridworld-stochastic
run X=1 $B --workload config3-dpw-gridworld
run MCTSB200_LANE_ORDER=0 $B --workload config3-dpw-gridworld
run X=1 $B --workload config3-dpw-gridworld --lanes 1
run X=1 $B --workload config3-dpw-gridworld --lanes 4
run X=1 timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload config4-dpw-mountaincar
cat gpurun_out/ab.log
