#!/bin/bash
# A/B of the successor-row request mechanism (MCTSB200_TOUCH) on config 4, plus the parity tests of the default build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "plugin or mountaincar or lanes or episodes" > gpurun_out/c4_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/c4_tests.log
for v in default touch0 touch1 touch3; do
  if [ $v = default ]; then unset MCTSB200_LIB; else export MCTSB200_LIB=$PWD/mcts.jl_b200/csrc/variants/lib$v.so; fi
  echo "== $v"
  python bench.py --workload config4-dpw-mountaincar --steps 2 --warmup 1 --no-cpu-baseline 2>gpurun_out/ab_$v.err | tee gpurun_out/ab_$v.json | python scripts/summarize_bench.py
done
