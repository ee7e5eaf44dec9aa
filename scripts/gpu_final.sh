# Round-end artefacts: GPU tests, default bench line, launch list and one full ncu capture of the same command.
set -x
R=${1:-r1}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_$R.log
python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke_$R.log
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -c 3000 gpurun_out/bench_$R.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ll_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"search_kernel|uct_fast_kernel|dpw_fast_kernel" -s 1 -c 1 -f -o gpurun_out/prof_full_$R python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$R.log 2>&1
rm -f gpurun_out/bench_other_$R.json
for W in config2-uct-gridworld-default config2-uct-gridworld-stochastic config3-dpw-gridworld; do python bench.py --workload $W --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/bench_other_$R.json; done
python bench.py --workload config4-dpw-mountaincar --no-cpu-baseline --steps 2 --warmup 1 2>/dev/null | tail -1 >> gpurun_out/bench_other_$R.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference_$R.json
