set -x
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
tail -c 1500 gpurun_out/bench_g$N.json; tail -3 gpurun_out/bench_g$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/root_parallel_nccl.py > gpurun_out/config5_g$N.json 2> gpurun_out/config5_g$N.err
tail -c 1000 gpurun_out/config5_g$N.json; tail -3 gpurun_out/config5_g$N.err
ENSEMBLE=scaled python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/root_parallel_nccl.py > gpurun_out/config5_scaled_g$N.json 2> gpurun_out/config5_scaled_g$N.err
tail -c 700 gpurun_out/config5_scaled_g$N.json
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline | tail -1 > gpurun_out/bench_g1_same_box.json
python scripts/root_parallel_nccl.py > gpurun_out/config5_g1.json 2>&1
tail -c 600 gpurun_out/config5_g1.json
