# round 2, call 49: HIndex side line, session-start build (tools/_variants/liblnr_base.so = commit bcd3a6f) against the final one
mkdir -p gpurun_out
export LNR_BENCH_NO_SMALL=1
cp linear_b200/csrc/liblnr_b200.so /tmp/liblnr_keep.so
for v in base cur; do
  if [ $v != cur ]; then cp tools/_variants/liblnr_$v.so linear_b200/csrc/liblnr_b200.so; else cp /tmp/liblnr_keep.so linear_b200/csrc/liblnr_b200.so; fi
  LNR_BENCH_INDEX=2 LNR_BENCH_GENOME=50000000 timeout 200 python bench.py --steps 4 --warmup 3 --batch-reads 16384 --no-cpu-baseline > gpurun_out/r2_bench49_hindex_$v.json 2> gpurun_out/r2_bench49_hindex_$v.err; echo "$v rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench49_hindex_$v.json'))
k=d['kernels_one_thread']
s=d['stage_cycles_last_batch']
print('$v', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), 'sort', round(k['k_hits_sort']['ms_per_launch'],2), {x: round(s[x]/1e9,1) for x in ('binning','sort_asc','run_filter','sort_x')}, s['n_tie_fallback'])
PY
done
cp /tmp/liblnr_keep.so linear_b200/csrc/liblnr_b200.so
