# round 2, call 24: bucket-scan masking restructured: parity subset + bench line; launch list of the repo's own kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q -x -k "apxmap_stages or against_unmodified or exhausted or N_runs or fuzz or random" 2>&1 | tail -5 > gpurun_out/r2_tests24.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench24.json 2> gpurun_out/r2_bench24.err; echo "bench rc=$?" >> gpurun_out/r2_tests24.log
LNR_BENCH_BATCH=32768 LNR_BENCH_NO_SMALL=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 3000 --csv --log-file gpurun_out/r2_launches_own.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_list3.log 2>&1
echo "list rc=$?" >> gpurun_out/r2_tests24.log
cat gpurun_out/r2_tests24.log
