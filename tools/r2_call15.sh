# round 2, call 15: hybrid -c 0 against the reference binary; HIndex (-i 2) side line of the bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hybrid.py -m gpu -q -x -k "c0" 2>&1 | tail -30 > gpurun_out/r2_tests15.log
LNR_BENCH_INDEX=2 timeout 900 python bench.py --steps 4 --warmup 3 --batch-reads 16384 > gpurun_out/r2_bench15_hindex.json 2> gpurun_out/r2_bench15_hindex.err; echo "hindex bench 3.1G rc=$?" >> gpurun_out/r2_tests15.log
if ! test -s gpurun_out/r2_bench15_hindex.json; then
  LNR_BENCH_INDEX=2 LNR_BENCH_GENOME=500000000 timeout 600 python bench.py --steps 4 --warmup 3 --batch-reads 16384 > gpurun_out/r2_bench15_hindex_500m.json 2> gpurun_out/r2_bench15_hindex_500m.err; echo "hindex bench 0.5G rc=$?" >> gpurun_out/r2_tests15.log
fi
tail -5 gpurun_out/r2_bench15_hindex.err >> gpurun_out/r2_tests15.log
cat gpurun_out/r2_tests15.log
