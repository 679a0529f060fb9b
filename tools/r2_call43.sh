# round 2, call 43: L2 fetch size of the genome feature loads of the window walk (k_map_extend): 256 B / 64 B against the default
mkdir -p gpurun_out
export LNR_BENCH_NO_SMALL=1
cp linear_b200/csrc/liblnr_b200.so /tmp/liblnr_keep.so
for v in cur walk256 walk64; do
  if [ $v != cur ]; then cp tools/_variants/liblnr_$v.so linear_b200/csrc/liblnr_b200.so; fi
  timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench43_$v.json 2> gpurun_out/r2_bench43_$v.err; echo "$v rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench43_$v.json'))
k=d['kernels_one_thread']
print('$v', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), 'extend', round(k['k_map_extend']['ms_per_launch'],3), 'blocks', round(k['k_hits_blocks']['ms_per_launch'],3), 'one-thread step', round(d['roofline']['whole_step']['ms_per_step_one_thread'],2))
PY
done
cp /tmp/liblnr_keep.so linear_b200/csrc/liblnr_b200.so
