# register budget of k_map_extend: rebuild with a min-CTAs launch bound (more resident warps, fewer registers), time the kernel
mkdir -p gpurun_out
cp linear_b200/csrc/liblnr_b200.so /tmp/liblnr_keep.so
for c in 10 12; do
  rm -f linear_b200/csrc/liblnr_b200.so
  LNR_NVCC_EXTRA="-DLNR_EXTEND_MIN_CTAS=$c" python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
  echo "extend min_ctas=$c"
  LNR_BENCH_NO_SMALL=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), round(d['kernels_one_thread']['k_map_extend']['ms_per_launch'],2), round(d['roofline']['whole_step']['ms_per_step_one_thread'],2))"
done
cp /tmp/liblnr_keep.so linear_b200/csrc/liblnr_b200.so
