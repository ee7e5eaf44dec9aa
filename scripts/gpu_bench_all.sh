# tests + bench lines for the main workloads (each bench line appended to gpurun_out/bench_all.log)
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
: > gpurun_out/bench_all.log
for W in config2-uct-gridworld-deterministic config2-uct-gridworld-stochastic config3-dpw-gridworld; do
  timeout 900 python bench.py --steps 3 --warmup 3 --workload $W --no-cpu-baseline 2>&1 | tail -1 >> gpurun_out/bench_all.log
done
timeout 900 python bench.py --steps 2 --warmup 1 --workload config4-dpw-mountaincar --no-cpu-baseline 2>&1 | tail -1 >> gpurun_out/bench_all.log
for L in 2 4; do timeout 600 python bench.py --steps 2 --warmup 1 --lanes $L --no-cpu-baseline 2>&1 | tail -1 >> gpurun_out/bench_all.log; done
for L in 1 2 8; do timeout 600 python bench.py --steps 2 --warmup 1 --lanes $L --workload config3-dpw-gridworld --no-cpu-baseline 2>&1 | tail -1 >> gpurun_out/bench_all.log; done
cat gpurun_out/bench_all.log | cut -c1-330
