# 8-GPU artefacts: weak-scaling bench (config 2 and config 4: 8 192 MountainCar roots = 1 024 per GPU), root-parallel config 5
N=${1:-8}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
tail -c 400 gpurun_out/bench_g$N.json; tail -2 gpurun_out/bench_g$N.err
$TR --master-port 29514 bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline --workload config4-dpw-mountaincar > gpurun_out/bench_c4_g$N.json 2> gpurun_out/bench_c4_g$N.err
tail -c 400 gpurun_out/bench_c4_g$N.json
$TR --master-port 29512 scripts/root_parallel_nccl.py > gpurun_out/config5_g$N.json 2> gpurun_out/config5_g$N.err
tail -c 600 gpurun_out/config5_g$N.json; tail -2 gpurun_out/config5_g$N.err
ENSEMBLE=scaled $TR --master-port 29513 scripts/root_parallel_nccl.py > gpurun_out/config5_scaled_g$N.json 2> gpurun_out/config5_scaled_g$N.err
tail -c 600 gpurun_out/config5_scaled_g$N.json
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline | tail -1 > gpurun_out/bench_g1_same_box.json
