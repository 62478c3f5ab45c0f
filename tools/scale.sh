# usage: tools/scale.sh NGPU "extra bench args" ["more args" ...] -- one torchrun bench per argument set; the JSON line
# of every run is appended to gpurun_out/scale_runs.jsonl
N=$1; shift
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], (d.get("parallelism") or {}).get("exchange"), d.get("exchange"), [(k["name"], round(k["ms_total"]/d["steps"],3)) for k in d.get("kernels",[])])'
mkdir -p gpurun_out
for a in "$@"; do
  echo "== $N GPUs: $a"
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-e2e $a 2>&1 | tail -1 | tee -a gpurun_out/scale_runs.jsonl | python -c "$P"
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e $a 2>&1 | tail -1 | tee -a gpurun_out/scale_runs.jsonl | python -c "$P"
  fi
done
