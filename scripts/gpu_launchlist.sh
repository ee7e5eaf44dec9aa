# usage: bash scripts/gpu_launchlist.sh <tag> [ENV=VAL ...] -- <bench args...>: ncu launch list (per-launch durations) of one bench run
TAG=$1; shift
ENVS=()
while [ "$1" != "--" ]; do ENVS+=("$1"); shift; done; shift
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline $*"
env "${ENVS[@]}" X=1 $CMD > gpurun_out/plain_ll_$TAG.log 2>&1 &&
env "${ENVS[@]}" X=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_ll_$TAG.log 2>&1
tail -n 1 gpurun_out/plain_ll_$TAG.log | cut -c1-200
