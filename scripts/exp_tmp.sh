for w in config4-dpw-mountaincar box-dpw-continuous-actions uct-mountaincar; do
  B="python bench.py --workload $w --no-extras --no-cpu-baseline --steps 3 --warmup 2"
  $B > gpurun_out/r2ad_${w}.json 2>/dev/null
  MCTSB200_CARVEOUT=100 $B > gpurun_out/r2ad_${w}_carve100.json 2>/dev/null
  MCTSB200_CARVEOUT=25 $B > gpurun_out/r2ad_${w}_carve25.json 2>/dev/null
done
for f in gpurun_out/r2ad_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -1)"; done
