B="python bench.py --workload config3-dpw-gridworld --no-extras --no-cpu-baseline --steps 5 --warmup 3"
V=$PWD/mcts.jl_b200/csrc/variants/libnonatab.so
for i in 1 2 3; do
$B > gpurun_out/r2w_c3_cur_$i.json 2>/dev/null
MCTSB200_LIB=$V $B > gpurun_out/r2w_c3_nonatab_$i.json 2>/dev/null
done
for f in gpurun_out/r2w_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -1)"; done
