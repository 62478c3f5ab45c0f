# usage: tools/scale.sh NGPU "extra bench args" ["more args" ...] -- one torchrun bench per argument set
N=$1; shift
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["config"].get("parallelism"), d.get("exchange"), [(k["name"], round(k["ms_total"]/d["steps"],3)) for k in d.get("kernels",[])])'
for a in "$@"; do
  echo "== $N GPUs: $a"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e $a 2>&1 | tail -1 | python -c "$P"
done
