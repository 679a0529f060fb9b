# round 2, call 39 / 45: launch list + one full ncu capture of the second step (after the same command exited 0 without ncu)
mkdir -p gpurun_out
export LNR_BENCH_BATCH=32768 LNR_BENCH_NO_SMALL=1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu45_plain.json 2> gpurun_out/r2_ncu45_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 3000 --csv --log-file gpurun_out/r2_launches45.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu45_list.log 2>&1
echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^(k_feat_reads|k_seed_count|k_seed_fill|k_hits_sort|k_hits_chain|k_hits_blocks|k_map_extend|k_map_finish)" -s 14 -c 14 -f -o gpurun_out/r2_full45 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu45_full.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/r2_ncu45_full.log | cut -c1-150
ls -la gpurun_out/r2_full45.ncu-rep
