# A/B of the TMA uses (GPU box): parity tests with everything on, then the cfg4 bench per ADMM_B200_TMA mask
# (0 none, 1 back-projector windows, 2 forward L2 prefetch, 4 forward y-dominant tile landing, 7 all).
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], [(k["name"], round(k["ms_total"]/k["launches"],3)) for k in d["kernels"][:8]])'
for m in ${TMA_MASKS:-0 1 2 4 7}; do
  echo "MASK $m"
  ADMM_B200_TMA=$m python bench.py --steps ${AB_STEPS:-6} --warmup 3 --no-cpu --no-e2e ${AB_ARGS:-} 2>gpurun_out/tma_ab_$m.err | tail -1 | python -c "$P" || tail -5 gpurun_out/tma_ab_$m.err
done
