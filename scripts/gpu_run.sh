# usage: bash scripts/gpu_run.sh <tag> <what...>   (on the GPU box, from the repo root; outputs under gpurun_out/)
#   tests            pytest -m gpu (whole suite)
#   tests:<expr>     pytest -m gpu -k <expr>
#   bench[:args]     python bench.py [args]            -> gpurun_out/<tag>_bench[_i].json
#   ref              python bench.py --impl reference   -> gpurun_out/<tag>_ref.json
#   launches[:args]  ncu launch list of bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras [args]
#   ncu:<name>:args  one ncu --set full capture of the search kernel of `bench.py args` -> gpurun_out/<tag>_<name>.ncu-rep + summaries
#   env:K=V          export K=V for the following items
#   lib:<path>       MCTSB200_LIB for the following items (a scripts/build_variant.sh build); lib: clears
TAG=$1; shift
mkdir -p gpurun_out
i=0
for what in "$@"; do
  kind=${what%%:*}; rest=${what#*:}; [ "$rest" = "$what" ] && rest=""
  case $kind in
    env) export "$rest";;
    lib) if [ -n "$rest" ]; then export MCTSB200_LIB=$rest; else unset MCTSB200_LIB; fi;;
    tests) if [ -n "$rest" ]; then python -m pytest tests -m gpu -x -q -k "$rest" > gpurun_out/${TAG}_tests_$i.full 2>&1; (grep -n "Fatal\|Segmentation\|^tests/.*::" gpurun_out/${TAG}_tests_$i.full | head -n 20; tail -n 25 gpurun_out/${TAG}_tests_$i.full) > gpurun_out/${TAG}_tests_$i.log; else python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 > gpurun_out/${TAG}_tests.log; fi;;
    bench) python bench.py $rest > gpurun_out/${TAG}_bench_$i.json 2> gpurun_out/${TAG}_bench_$i.err;;
    ref) python bench.py --impl reference $rest > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err;;
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras $rest > gpurun_out/${TAG}_launches.log 2>&1;;
    ncu) name=${rest%%:*}; args=${rest#*:}
         python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras $args > gpurun_out/${TAG}_${name}_plain.json 2>&1 &&
         ncu --set full --clock-control none --import-source on -k regex:"search_kernel|uct_fast_kernel|dpw_fast_kernel" -s 1 -c 1 -f -o gpurun_out/${TAG}_${name} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras $args > gpurun_out/${TAG}_${name}_ncu.log 2>&1
         python scripts/ncu_summary.py gpurun_out/${TAG}_${name}.ncu-rep > gpurun_out/${TAG}_${name}_ncu_summary.txt 2>&1;;
    *) echo "unknown item $what";;
  esac
  i=$((i+1))
done
ls -la gpurun_out | tail -n 30
