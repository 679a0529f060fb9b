# round 2, call 28: (1) parity of the two-strand k_feat_reads and of the 1-rank form of the sharded HIndex build;
# (2) same-box A/B of prebuilt library variants (tools/_variants/): old k_feat_reads, k_map_extend at 12 / 16 CTAs per SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded_build.py -m gpu -q -x -k "feature or hindex or apxmap_stages or against_unmodified or sharded or packed" 2>&1 | tail -5 > gpurun_out/r2_tests28.log
cat gpurun_out/r2_tests28.log
export LNR_BENCH_NO_SMALL=1
cp linear_b200/csrc/liblnr_b200.so /tmp/liblnr_keep.so
for v in cur oldfeat mc12 mc16; do
  if [ $v != cur ]; then cp tools/_variants/liblnr_$v.so linear_b200/csrc/liblnr_b200.so; fi
  timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench28_$v.json 2> gpurun_out/r2_bench28_$v.err; echo "$v rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench28_$v.json'))
k=d['kernels_one_thread']
print('$v', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), 'feat', round(k['k_feat_reads']['ms_per_launch'],3), 'extend', round(k['k_map_extend']['ms_per_launch'],3), 'one-thread step', round(d['roofline']['whole_step']['ms_per_step_one_thread'],2))
PY
done
cp /tmp/liblnr_keep.so linear_b200/csrc/liblnr_b200.so
