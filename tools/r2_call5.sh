# round 2, call 5: -f 1 on the GPU, index build v3 (place -> rank, pooled temporaries): all GPU tests, bench, ncu of the build
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_tests5.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc=$?" >> gpurun_out/r2_tests5.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^(k_idx_emit|k_idx_partcount|k_idx_part|k_idx_count|k_idx_place|k_idx_rank|k_idx_dirx)" -c 10 -f -o gpurun_out/r2_idx4_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_idx4.log 2>&1
cat gpurun_out/r2_tests5.log
