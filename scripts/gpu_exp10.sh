set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
: > gpurun_out/exp10.log
B="timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { echo "## $*" >> gpurun_out/exp10.log; env "$@" 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/exp10.log; }
run X=1 $B
run X=1 $B --workload config2-uct-gridworld-stochastic
run X=1 $B --workload config3-dpw-gridworld
run MCTSB200_SPREAD=0 $B --workload config3-dpw-gridworld
run X=1 $B --workload config2-uct-gridworld-stochastic --roots 16384
run MCTSB200_SPREAD=0 $B --workload config2-uct-gridworld-stochastic --roots 16384
run X=1 timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload config4-dpw-mountaincar
cat gpurun_out/exp10.log
python scripts/root_parallel_nccl.py 2>&1 | tail -1 | cut -c1-400
MCTSB200_SPREAD=0 python scripts/root_parallel_nccl.py 2>&1 | tail -1 | cut -c1-400
