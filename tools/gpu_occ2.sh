# occupancy experiment for k_map_hits: rebuild with a min-CTAs launch bound, then time at matching CTAs/SM
for c in 3 4 5 6; do
  rm -f linear_b200/csrc/liblnr_b200.so
  LNR_NVCC_EXTRA="-DLNR_HITS_MIN_CTAS=$c" python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
  echo "min_ctas=$c"
  LNR_BENCH_BATCH=32768 LNR_MAP_CTAS_PER_SM=$c python bench.py --steps 3 --warmup 2 --no-cpu-baseline --streams 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), round(d['kernels']['k_map_hits']['ms_per_launch'],2))"
done
