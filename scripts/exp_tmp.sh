V=$PWD/mcts.jl_b200/csrc/variants/libdpwred.so
MCTSB200_LIB=$V python -m pytest tests -m gpu -x -q -k "dpw or DPW or golden" > gpurun_out/r2o_tests.full 2>&1; tail -n 3 gpurun_out/r2o_tests.full
B="python bench.py --workload config3-dpw-gridworld --no-extras --no-cpu-baseline --steps 3 --warmup 3"
$B > gpurun_out/r2o_c3_base.json 2>/dev/null
MCTSB200_LIB=$V $B > gpurun_out/r2o_c3_red.json 2>/dev/null
$B --roots 65536 > gpurun_out/r2o_c3_64k_base.json 2>/dev/null
MCTSB200_LIB=$V $B --roots 65536 > gpurun_out/r2o_c3_64k_red.json 2>/dev/null
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/r2o_*.json
