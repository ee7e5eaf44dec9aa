# debug experiment: rebuild with phase timing, run short benches (args separated by ';;'), rebuild clean
set -x
touch mcts.jl_b200/csrc/*.cuh; make -C mcts.jl_b200/csrc -j16 EXTRA=-DMCTSB200_PHASE_TIMING > /dev/null 2>&1
: > gpurun_out/phase_timing.log
IFS=';' read -ra RUNS <<< "$*"
for r in "${RUNS[@]}"; do
  [ -z "$r" ] && continue
  echo "## $r" >> gpurun_out/phase_timing.log
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline $r 2>&1 | grep -E "^PT|metric" | sort | uniq | head -12 | cut -c1-330 >> gpurun_out/phase_timing.log
done
touch mcts.jl_b200/csrc/*.cuh; make -C mcts.jl_b200/csrc -j16 > /dev/null 2>&1
cat gpurun_out/phase_timing.log
