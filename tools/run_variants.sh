# usage (GPU box): bash tools/run_variants.sh v1 v2 ...   -- benches build_variants/lib_<v>.so one after the other
for v in "$@"; do
  cp build_variants/lib_$v.so pysurfinv_b200/libsurfdisp_b200.so
  echo "== $v"
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['love']['ms_per_step'])"
done
