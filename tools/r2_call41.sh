# round 2, call 41: explicit L2 prefetch size on the random hs record loads of k_seed_fill (LDG.E.LTC64B / 128B / 256B) against
# the plain load: which fetch granularity does the default have?
mkdir -p gpurun_out
export LNR_BENCH_NO_SMALL=1
cp linear_b200/csrc/liblnr_b200.so /tmp/liblnr_keep.so
for v in prev l2_64 l2_128 l2_256; do
  cp tools/_variants/liblnr_$v.so linear_b200/csrc/liblnr_b200.so
  timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench41_$v.json 2> gpurun_out/r2_bench41_$v.err; echo "$v rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench41_$v.json'))
k=d['kernels_one_thread']
print('$v', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), 'count', round(k['k_seed_count']['ms_per_launch'],3), 'fill', round(k['k_seed_fill']['ms_per_launch'],3), 'one-thread step', round(d['roofline']['whole_step']['ms_per_step_one_thread'],2))
PY
done
cp /tmp/liblnr_keep.so linear_b200/csrc/liblnr_b200.so
