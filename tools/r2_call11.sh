# round 2, call 11: occupancy experiments (rebuilds on the box), then the chunked-CLI / FASTQ tests with the shipped build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cli_apf.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_tests11.log
cp linear_b200/csrc/liblnr_b200.so /tmp/lib_ship.so
run() { LNR_BENCH_BATCH=32768 "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline --streams 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels_one_thread']; print(round(d['value']), round(d['ms_per_step'],2), {n:round(k[n]['ms_per_launch'],3) for n in ('k_hits_sort','k_hits_chain','k_hits_blocks','k_map_extend','k_map_finish')})"; }
echo "shipped" >> gpurun_out/r2_tests11.log; run env >> gpurun_out/r2_tests11.log 2>&1
for c in 8 10; do
  rm -f linear_b200/csrc/liblnr_b200.so
  LNR_NVCC_EXTRA="-DLNR_EXTEND_MIN_CTAS=$c" python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
  echo "extend min_ctas=$c" >> gpurun_out/r2_tests11.log; run env >> gpurun_out/r2_tests11.log 2>&1
done
for c in 8 10; do
  rm -f linear_b200/csrc/liblnr_b200.so
  LNR_NVCC_EXTRA="-DLNR_CHAIN_MIN_CTAS=$c -DLNR_BLOCKS_MIN_CTAS=$c" python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
  echo "chain/blocks min_ctas=$c" >> gpurun_out/r2_tests11.log; run env LNR_CHAIN_CTAS_PER_SM=$c LNR_BLOCKS_CTAS_PER_SM=$c >> gpurun_out/r2_tests11.log 2>&1
done
cp /tmp/lib_ship.so linear_b200/csrc/liblnr_b200.so
cat gpurun_out/r2_tests11.log
