# run one gpurun command, retrying while the pod answers "no slot" (exit 3, nothing charged)
# usage: [GPUS=N] tools/gpu_retry.sh <timeout-seconds> '<command>'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$1" -- "$2"; rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 90
done
echo "gpu_retry rc=$rc"
