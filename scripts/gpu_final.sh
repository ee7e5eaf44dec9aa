# Round-end artefacts on the GPU box: GPU tests, smoke, the default bench line, the ncu launch list of the same command, one full
# ncu capture per workload family, the reference arm.   usage: bash scripts/gpu_final.sh <tag>
set -x
R=${1:-r2}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${R}_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -6 | tee gpurun_out/${R}_smoke.log
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${R}_ncu_ll.log 2>&1
cap() {  # <name> <bench args...>: one full capture of the search kernel of a workload (after the same command ran clean without ncu)
  N=$1; shift
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras "$@" > gpurun_out/${R}_${N}_plain.json 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"search_kernel|uct_fast_kernel|dpw_fast_kernel" -s 1 -c 1 -f -o gpurun_out/${R}_${N} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras "$@" > gpurun_out/${R}_${N}_ncu.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${R}_${N}.ncu-rep > gpurun_out/${R}_${N}_ncu_summary.txt 2>&1
}
cap uct_default
cap uct_det --workload config2-uct-gridworld-deterministic
cap uct_stoch --workload config2-uct-gridworld-stochastic
cap dpw_c3 --workload config3-dpw-gridworld
cap mc_c4 --workload config4-dpw-mountaincar --iterations 2000
cap c5 --workload config5-root-parallel-uct-gridworld --same-root 1,1
ls -la gpurun_out | tail -n 40
