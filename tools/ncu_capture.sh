# ncu evidence for profiles/ (one GPU).  Each ncu pass runs only after the same command exited 0 without ncu.
# usage: tools/ncu_capture.sh <tag> [launches] [proj] [stream]
# One warm outer iteration of the default bench (cfg4, 1 sweep x 2 CG per solve, a14 rule on: 3 solves, residual carried
# between them) is 35 launches: rhs0, back<2>, 3 x (fwd, reduce, back<1>, axpy, fwd(fused), reduce, back<1>, axpy,
# cg_update, tv), sino_resid, edge, finalize.  The launch list covers ~3 iterations; profiles/make_profiles.py cuts one
# complete iteration out of it (between two rhs0 launches).
set -e
TAG=$1; shift
export ADMM_B200_NOGRAPH=1     # eager launches, so that -s / -c count kernels of the bench loop itself
CMD="python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu --no-e2e --no-profile"
$CMD > gpurun_out/plain.log 2>&1
for what in "$@"; do
  if [ "$what" = "launches" ]; then
    ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${TAG}_launches_cfg4.csv $CMD > gpurun_out/ncu1.log 2>&1
  fi
  if [ "$what" = "proj" ]; then
    # projector launches of the 4th iteration in order: back<2>, fwd, back<1>, fwd (CG-fused), back<1> (13 per iteration)
    ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"fwd_strip|back_tile_kernelILi[12]" -s 41 -c 5 -f -o gpurun_out/${TAG}_cfg4_proj $CMD > gpurun_out/ncu2.log 2>&1
    ncu -i gpurun_out/${TAG}_cfg4_proj.ncu-rep --page raw --csv > gpurun_out/${TAG}_cfg4_proj_raw.csv
  fi
  if [ "$what" = "stream" ]; then
    ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"tv_fused|edge_kernel|rhs0_kernel|cg_update" -s 24 -c 8 -f -o gpurun_out/${TAG}_cfg4_stream $CMD > gpurun_out/ncu3.log 2>&1
    ncu -i gpurun_out/${TAG}_cfg4_stream.ncu-rep --page raw --csv > gpurun_out/${TAG}_cfg4_stream_raw.csv
  fi
done
ls -la gpurun_out/ | tail -12
