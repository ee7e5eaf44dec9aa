# debug experiment: rebuild with phase timing, run one short bench, rebuild clean
set -x
make -C mcts.jl_b200/csrc clean >/dev/null; make -C mcts.jl_b200/csrc -j16 EXTRA=-DMCTSB200_PHASE_TIMING > /dev/null 2>&1
python bench.py --steps 1 --warmup 0 --no-cpu-baseline $* 2>&1 | grep -E "^PT|metric" | head -40 | cut -c1-250 > gpurun_out/phase_timing.log
make -C mcts.jl_b200/csrc clean >/dev/null; make -C mcts.jl_b200/csrc -j16 > /dev/null 2>&1
cat gpurun_out/phase_timing.log
