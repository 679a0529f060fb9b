# round 2, call 8: source-level ncu of the hit-stage / extension / finish kernels, small-block call anatomy
mkdir -p gpurun_out
timeout 600 python tools/gpu_small_blocks.py > gpurun_out/r2_small_blocks.log 2>&1
LNR_BENCH_BATCH=32768 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^(k_hits_sort|k_hits_chain|k_hits_blocks|k_map_extend|k_map_finish|k_seed_count|k_seed_fill|k_feat_reads)$" -s 8 -c 8 -f -o gpurun_out/r2_hits_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_hits.log 2>&1
tail -2 gpurun_out/r2_ncu_hits.log | cut -c1-200
cat gpurun_out/r2_small_blocks.log | tail -40
