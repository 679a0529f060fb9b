# k_map_hits: register cap (launch bound) x resident CTAs/SM -> queue drain time and kernel span
for c in 3 4 5; do
  rm -f linear_b200/csrc/liblnr_b200.so
  LNR_NVCC_EXTRA="-DLNR_HITS_MIN_CTAS=$c" python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
  echo "min_ctas=$c"
  LNR_BENCH_BATCH=32768 LNR_MAP_CTAS_PER_SM=$c LNR_LONGEST_PROFILE=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 2>gpurun_out/err.log >/dev/null
  grep "lnr tail" gpurun_out/err.log | tail -16 | head -3 | cut -c1-250
done
