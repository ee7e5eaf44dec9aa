# usage: bash scripts/gpu_ncu2.sh <tag> [ENV=VAL ...] -- <bench args...>: one full ncu capture of the search kernel
set -x
TAG=$1; shift
ENVS=()
while [ "$1" != "--" ]; do ENVS+=("$1"); shift; done; shift
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline $*"
env "${ENVS[@]}" X=1 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
env "${ENVS[@]}" X=1 ncu --set full --clock-control none --import-source on -k regex:"search_kernel|uct_fast_kernel|dpw_fast_kernel" -s 1 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_full_$TAG.log
