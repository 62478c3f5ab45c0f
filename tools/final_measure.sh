# Round-end measurement set on ONE B200 (GPU box): the default bench line with every leg, the reference arm, and one line
# per BASELINE config; outputs under gpurun_out/ (copied into profiles/ by hand afterwards).
# usage: tools/final_measure.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err; tail -c 600 gpurun_out/${TAG}_bench_cfg4.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
for c in cfg1 cfg2 cfg3 cfg5; do
  python bench.py --config $c --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/*_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    e2e = (d.get("e2e") or {}).get("value")
    print(f, d.get("impl", "ours"), "ms/step", d.get("ms_per_step"), "value", d.get("value"), "e2e", e2e,
          "iter roofline", (d.get("iteration_roofline") or {}).get("frac"), "top", (d.get("roofline") or {}).get("kernel"),
          (d.get("roofline") or {}).get("frac"), "launch", d.get("launch_mode"))
PY
