# ncu evidence for profiles/ (one GPU).  Run only after the same command has exited 0 without ncu.
# usage: tools/ncu_capture.sh <tag> [launches] [proj] [stream]
set -e
TAG=$1; shift
CMD="python bench.py --config cfg4 --steps 1 --no-cpu --no-e2e --no-profile"
$CMD > gpurun_out/plain.log 2>&1
for what in "$@"; do
  if [ "$what" = "launches" ]; then
    ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 44 --csv --log-file gpurun_out/${TAG}_launches_cfg4.csv $CMD > gpurun_out/ncu1.log 2>&1
  fi
  if [ "$what" = "proj" ]; then
    # filtered launch order in a warm iteration: fwd (plain), back<2>, [fwd (plain) back<1>], fwd (CG-fused), back<1>, ...
    ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"fwd_strip|back_tile_kernelILi1" -s 34 -c 3 -f -o gpurun_out/${TAG}_cfg4_proj $CMD > gpurun_out/ncu2.log 2>&1
    ncu -i gpurun_out/${TAG}_cfg4_proj.ncu-rep --page raw --csv > gpurun_out/${TAG}_cfg4_proj_raw.csv
  fi
  if [ "$what" = "stream" ]; then
    ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"tv_fused|edge_kernel|rhs0_kernel|cg_update" -s 12 -c 4 -f -o gpurun_out/${TAG}_cfg4_stream $CMD > gpurun_out/ncu3.log 2>&1
    ncu -i gpurun_out/${TAG}_cfg4_stream.ncu-rep --page raw --csv > gpurun_out/${TAG}_cfg4_stream_raw.csv
  fi
done
ls -la gpurun_out/
