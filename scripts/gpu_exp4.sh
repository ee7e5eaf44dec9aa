set -x
: > gpurun_out/exp4.log
B="timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { echo "## $*" >> gpurun_out/exp4.log; env "$@" 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/exp4.log; }
for i in 1 2; do
run MCTSB200_LANE_ORDER=1 $B
run MCTSB200_LANE_ORDER=1 BENCH_NO_SAMPLER=1 $B
run X=1 $B --workload config2-uct-gridworld-stochastic
run BENCH_NO_SAMPLER=1 $B --workload config2-uct-gridworld-stochastic
done
cat gpurun_out/exp4.log
