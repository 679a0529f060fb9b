# round 2, call 19: new tests (index save/load, HIndex at 200 Mbase, -c 0), HIndex side line at 50 Mbase, main bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "c0 or save_load or 200_mbase" 2>&1 | tail -15 > gpurun_out/r2_tests19.log
LNR_BENCH_INDEX=2 LNR_BENCH_GENOME=50000000 LNR_BENCH_NO_SMALL=1 timeout 200 python bench.py --steps 4 --warmup 3 --batch-reads 16384 --no-cpu-baseline > gpurun_out/r2_bench19_hindex_50m.json 2> gpurun_out/r2_bench19_hindex_50m.err; echo "hindex 50M bench rc=$?" >> gpurun_out/r2_tests19.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err; echo "bench rc=$?" >> gpurun_out/r2_tests19.log
cat gpurun_out/r2_tests19.log
