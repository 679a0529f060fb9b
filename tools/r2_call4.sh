# round 2, call 4: index build v2 (sampled splitters, L2-local count/place/sort): tests, bench, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py tests/test_gpu_sharded_build.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_tests4.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?" >> gpurun_out/r2_tests4.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^(k_idx_emit|k_idx_partcount|k_idx_part|k_idx_count|k_idx_place|k_idx_sort_part)" -c 12 -f -o gpurun_out/r2_idx3_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_idx3.log 2>&1
cat gpurun_out/r2_tests4.log
