# usage: bash scripts/gpurun_retry.sh [--gpus N] <timeout> <command...>  -- retries a refused / busy gpurun call every 3 minutes (here, not on the box)
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $GP --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  if grep -q "status=ok\|status=failed\|status=timeout" /tmp/gpurun_last.log; then break; fi
  sleep 180
done
tail -n 6 /tmp/gpurun_last.log
