# round 2, call 9: source-level ncu of the hit-stage / extension / finish kernels (second step, primary pass), tests of the
# batch-sized workspace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_tests9.log
LNR_BENCH_BATCH=32768 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^(k_hits_sort|k_hits_chain|k_hits_blocks|k_map_extend|k_map_finish)$" -s 9 -c 5 -f -o gpurun_out/r2_hits_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_hits.log 2>&1
tail -2 gpurun_out/r2_ncu_hits.log | cut -c1-200 >> gpurun_out/r2_tests9.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc=$?" >> gpurun_out/r2_tests9.log
cat gpurun_out/r2_tests9.log
