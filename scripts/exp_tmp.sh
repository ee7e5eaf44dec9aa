V=$PWD/mcts.jl_b200/csrc/variants/libtab512.so
for w in config2-uct-gridworld-default config2-uct-gridworld-stochastic config2-uct-gridworld-deterministic config3-dpw-gridworld config5-root-parallel-uct-gridworld; do
  B="python bench.py --workload $w --no-extras --no-cpu-baseline --steps 4 --warmup 3"
  $B > gpurun_out/r2aa_${w}_base.json 2>/dev/null
  MCTSB200_LIB=$V $B > gpurun_out/r2aa_${w}_tab512.json 2>/dev/null
done
MCTSB200_LIB=$V python -m pytest tests -m gpu -x -q -k "uct or dpw or visit_counts" > gpurun_out/r2aa_tests.full 2>&1; tail -n 3 gpurun_out/r2aa_tests.full
for f in gpurun_out/r2aa_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -1)"; done
