#!/bin/bash
bash scripts/gpu_quick.sh
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload config4-dpw-mountaincar 2>&1 | tail -1 | python scripts/summarize_bench.py
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload config4-dpw-mountaincar --roots 8192 2>&1 | tail -1 | python scripts/summarize_bench.py
