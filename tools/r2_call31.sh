# round 2, call 31: binning filter over a candidate list (2 passes over A instead of 4) + prefetching radix passes:
# parity (stages, fuzz, unmodified reference), then A/B against the previous build (tools/_variants/liblnr_feat1buf.so)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_tests31.log
cat gpurun_out/r2_tests31.log
export LNR_BENCH_NO_SMALL=1
cp linear_b200/csrc/liblnr_b200.so /tmp/liblnr_keep.so
for v in cur feat1buf; do
  if [ $v != cur ]; then cp tools/_variants/liblnr_$v.so linear_b200/csrc/liblnr_b200.so; fi
  timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench31_$v.json 2> gpurun_out/r2_bench31_$v.err; echo "$v rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench31_$v.json'))
k=d['kernels_one_thread']
s=d['stage_cycles_last_batch']
print('$v', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), 'sort', round(k['k_hits_sort']['ms_per_launch'],3), 'chain', round(k['k_hits_chain']['ms_per_launch'],3), 'blocks', round(k['k_hits_blocks']['ms_per_launch'],3), 'one-thread step', round(d['roofline']['whole_step']['ms_per_step_one_thread'],2), {x: round(s[x]/1e9,2) for x in ('binning','sort_asc','run_filter','sort_x','hit_blocks')})
PY
done
cp /tmp/liblnr_keep.so linear_b200/csrc/liblnr_b200.so
