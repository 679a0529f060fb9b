# register budget of the chain / blocks section kernels: rebuild with other launch bounds, run at matching residency
for c in 4; do
  rm -f linear_b200/csrc/liblnr_b200.so
  LNR_NVCC_EXTRA="-DLNR_CHAIN_MIN_CTAS=$c -DLNR_BLOCKS_MIN_CTAS=$c" python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
  echo "chain/blocks min_ctas=$c"
  LNR_BENCH_BATCH=32768 LNR_CHAIN_CTAS_PER_SM=$c LNR_BLOCKS_CTAS_PER_SM=$c python bench.py --steps 3 --warmup 2 --no-cpu-baseline --streams 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), {k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items() if k.startswith('k_hits')})"
  LNR_CHAIN_CTAS_PER_SM=$c LNR_BLOCKS_CTAS_PER_SM=$c python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('4 streams', round(d['value']), round(d['e2e']['value']))"
done
