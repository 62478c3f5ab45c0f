# ncu evidence for profiles/ (one GPU).  Run only after the same command has exited 0 without ncu.
# usage: tools/ncu_capture.sh [launches] [full]
set -e
CMD="python bench.py --config cfg4 --steps 1 --no-cpu --no-e2e --no-profile"
$CMD > gpurun_out/plain.log 2>&1
if [ "$1" = "launches" ] || [ "$2" = "launches" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r1_launches_cfg4.csv $CMD > gpurun_out/ncu1.log 2>&1
fi
if [ "$1" = "full" ] || [ "$2" = "full" ]; then
  # filtered launch order in a warm iteration: fwd (plain), back<1>, fwd (CG-fused), back<1>, ...
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"fwd_strip|back_tile_kernelILi1" -s 33 -c 3 -f -o gpurun_out/r1_cfg4_proj $CMD > gpurun_out/ncu2.log 2>&1
  ncu -i gpurun_out/r1_cfg4_proj.ncu-rep --page raw --csv > gpurun_out/r1_cfg4_proj_raw.csv
fi
ls -la gpurun_out/
