#!/bin/bash
# GPU check of the run-time plugin path: the plugin tests first (short), then the whole -m gpu suite.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k plugin > gpurun_out/plugin_tests.log 2>&1
echo "plugin tests exit $?" | tee -a gpurun_out/plugin_tests.log
tail -30 gpurun_out/plugin_tests.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log
