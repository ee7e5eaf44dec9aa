#!/bin/bash
# quick GPU check: parity suite, then the three bench workloads (no CPU leg)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/quick.log
for w in config2-uct-gridworld-deterministic config2-uct-gridworld-stochastic config3-dpw-gridworld; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $w 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l); continue
    print(d['config']['workload'], 'ms', round(d['ms_per_step'],2), 'Msim/s', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])
" >> gpurun_out/quick.log
done
cat gpurun_out/quick.log
