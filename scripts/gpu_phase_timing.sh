# debug experiment: rebuild with phase timing, run short benches (args separated by ';'), rebuild clean
set -x
touch mcts.jl_b200/csrc/*.cuh; make -C mcts.jl_b200/csrc -j16 EXTRA="-DMCTSB200_PHASE_TIMING $PT_EXTRA" > /dev/null 2>&1
i=0
IFS=';' read -ra RUNS <<< "$*"
for r in "${RUNS[@]}"; do
  [ -z "$r" ] && continue
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline $r 2>&1 | grep -E "^PT|metric" | cut -c1-400 > gpurun_out/pt_$i.log
  i=$((i+1))
done
touch mcts.jl_b200/csrc/*.cuh; make -C mcts.jl_b200/csrc -j16 > /dev/null 2>&1
wc -l gpurun_out/pt_*.log
