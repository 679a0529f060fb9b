# round 2, call 3: new index build (emit/partition/place) + sharded build in the C ABI (1 rank): tests, bench, ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_tests3.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?" >> gpurun_out/r2_tests3.log
ncu --set full --clock-control none --import-source on -k regex:"^(k_idx_emit|k_idx_part|k_idx_place|k_idx_sort_part|k_scan_)" -c 9 -f -o gpurun_out/r2_idx2_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_idx2.log 2>&1
LNR_BENCH_BATCH=32768 ncu --set full --clock-control none --import-source on -k regex:"^(k_seed_count|k_seed_fill)$" -s 4 -c 2 -f -o gpurun_out/r2_seed_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_seed.log 2>&1
tail -2 gpurun_out/r2_ncu_seed.log | cut -c1-200 >> gpurun_out/r2_tests3.log
cat gpurun_out/r2_tests3.log
