# round 2, call 42: 64-byte L2 fetch hint on the hs record loads (k_seed_fill) and on the lookup entries (k_seed_count): parity subset,
# then the bench against a build whose lookup entries use the plain load
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q -x -k "apxmap_stages or against_unmodified or exhausted or N_runs or fuzz or random or packed" 2>&1 | tail -5 > gpurun_out/r2_tests42.log
cat gpurun_out/r2_tests42.log
export LNR_BENCH_NO_SMALL=1
cp linear_b200/csrc/liblnr_b200.so /tmp/liblnr_keep.so
for v in cur dirxplain; do
  if [ $v != cur ]; then cp tools/_variants/liblnr_$v.so linear_b200/csrc/liblnr_b200.so; fi
  timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench42_$v.json 2> gpurun_out/r2_bench42_$v.err; echo "$v rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench42_$v.json'))
k=d['kernels_one_thread']
print('$v', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), 'count', round(k['k_seed_count']['ms_per_launch'],3), 'fill', round(k['k_seed_fill']['ms_per_launch'],3), 'one-thread step', round(d['roofline']['whole_step']['ms_per_step_one_thread'],2))
PY
done
cp /tmp/liblnr_keep.so linear_b200/csrc/liblnr_b200.so
