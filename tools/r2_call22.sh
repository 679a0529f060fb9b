# round 2, call 22: launch list + one full capture of the per-step kernels on the final code (after the same command ran plain)
mkdir -p gpurun_out
export LNR_BENCH_BATCH=32768 LNR_BENCH_NO_SMALL=1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_list2.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"^(k_feat_reads|k_seed_count|k_seed_fill|k_hits_sort|k_hits_chain|k_hits_blocks|k_map_extend|k_map_finish)$" -s 14 -c 14 -f -o gpurun_out/r2_final_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_full_final.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/r2_ncu_full_final.log | cut -c1-160
