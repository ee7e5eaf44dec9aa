set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
: > gpurun_out/ab3.log
B="timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline"
run() { echo "## $*" >> gpurun_out/ab3.log; env "$@" 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/ab3.log; }
run X=1 $B
run MCTSB200_LANE_ORDER=0 $B
run X=1 $B --workload config2-uct-gridworld-stochastic
run MCTSB200_LANE_ORDER=0 $B --workload config2-uct-gridworld-stochastic
run X=1 $B --same-root 1,1 --roots 4736
run X=1 $B --roots 262144
cat gpurun_out/ab3.log
bash scripts/gpu_phase_timing.sh "--same-root 1,1 --roots 4736;--same-root 5,5 --roots 4736; ;--workload config2-uct-gridworld-stochastic"
