# A/B helper (GPU box): bench the current build, then each sed-variant given as arguments ("file:::sed-expr").
# Every variant is applied to a backup-restored copy of the file, so the tree is left exactly as it was found.
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e ${AB_ARGS:-}"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], [(k["name"], round(k["ms_total"]/k["launches"],3)) for k in d["kernels"][:7]])'
echo BASE; $B 2>&1 | tail -1 | python -c "$P"
for v in "$@"; do
  f="${v%%:::*}"; e="${v##*:::}"
  cp "$f" "$f.ab_backup"
  sed -i "$e" "$f"
  python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1 || echo BUILD FAILED
  echo "VARIANT $e"; $B 2>&1 | tail -1 | python -c "$P"
  mv "$f.ab_backup" "$f"
done
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
