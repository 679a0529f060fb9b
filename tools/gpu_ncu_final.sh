# launch list + one full capture of the second step (run only after the same command exited 0 without ncu)
export LNR_BENCH_BATCH=32768
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none -k regex:"^(k_feat_reads|k_seed_count|k_seed_fill|k_hits_sort|k_hits_chain|k_hits_blocks|k_map_extend|k_map_finish)" -s 14 -c 14 -f -o gpurun_out/full_v4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/ncu_full4.log 2>&1
tail -2 gpurun_out/ncu_full4.log | cut -c1-150
