set -x
: > gpurun_out/exp9.log
B="timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload config4-dpw-mountaincar"
run() { echo "## $*" >> gpurun_out/exp9.log; env "$@" 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/exp9.log; }
for L in 1 4 8 16 32; do run X=1 $B --lanes $L --iterations 2000; done
run MCTSB200_PATH_SMEM=0 $B --lanes 32 --iterations 2000
cat gpurun_out/exp9.log
