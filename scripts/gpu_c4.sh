#!/bin/bash
# config 4 (MountainCar DPW, tile kernel): GPU parity tests of the models that run on the tile kernel, bench at 1024 / 8192 roots,
# optional phase timing (PT=1)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "plugin or mountaincar or lanes or episodes" > gpurun_out/c4_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/c4_tests.log
python bench.py --workload config4-dpw-mountaincar --steps 2 --warmup 1 --no-cpu-baseline 2>gpurun_out/c4_1024.err | tee gpurun_out/c4_1024.json | python scripts/summarize_bench.py 2>/dev/null || tail -c 600 gpurun_out/c4_1024.json
python bench.py --workload config4-dpw-mountaincar --roots 8192 --steps 1 --warmup 1 --no-cpu-baseline 2>gpurun_out/c4_8192.err | tee gpurun_out/c4_8192.json | python scripts/summarize_bench.py 2>/dev/null || tail -c 600 gpurun_out/c4_8192.json
if [ -n "$PT" ]; then
  bash scripts/gpu_phase_timing.sh "--workload config4-dpw-mountaincar --iterations 4000" > gpurun_out/pt_run.log 2>&1; head -6 gpurun_out/pt_0.log
fi
