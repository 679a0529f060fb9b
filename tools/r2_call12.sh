# round 2, call 12: SAM* text, streaming-samples hybrid, full GPU suite on the current build, bench + ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_tests12.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; echo "bench rc=$?" >> gpurun_out/r2_tests12.log
LNR_BENCH_BATCH=32768 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_list.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/r2_tests12.log
cat gpurun_out/r2_tests12.log
