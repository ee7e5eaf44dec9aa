set -x
: > gpurun_out/exp1.log
B="timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
run() { echo "## $*" >> gpurun_out/exp1.log; env "$@" 2>&1 | tail -1 | python scripts/summarize_bench.py >> gpurun_out/exp1.log; }
for R in 1,1 5,5 4,2 9,2; do
run X=1 $B --same-root $R --roots 4736
run MCTSB200_PATH_SMEM=0 $B --same-root $R --roots 4736
run X=1 $B --same-root $R --roots 65536
done
run X=1 $B --roots 4736
run X=1 $B --roots 16384
run X=1 $B --roots 32768
run MCTSB200_LANE_ORDER=0 $B --roots 4736
run MCTSB200_LANE_ORDER=0 $B --roots 16384
run MCTSB200_LANE_ORDER=0 $B --roots 32768
cat gpurun_out/exp1.log
