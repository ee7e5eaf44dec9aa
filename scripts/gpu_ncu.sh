# usage: bash scripts/gpu_ncu.sh <tag> <bench args...>   -- launch list + one full capture of the search kernel
set -x
TAG=$1; shift
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"search_kernel|uct_fast_kernel|dpw_fast_kernel" -s 1 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/plain_$TAG.log gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/
