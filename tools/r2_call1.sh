# round 2, call 1: GPU tests, the bench with the parity object, L2 fetch granularity experiment, index-build ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_tests1.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?" >> gpurun_out/r2_tests1.log
python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/r2_ref1.json 2> gpurun_out/r2_ref1.err; echo "ref rc=$?" >> gpurun_out/r2_tests1.log
for g in 32 128; do
  LNR_L2_FETCH=$g LNR_BENCH_BATCH=32768 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/r2_l2fetch_$g.json 2> gpurun_out/r2_l2fetch_$g.err
done
LNR_BENCH_BATCH=32768 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/r2_l2fetch_default.json 2> gpurun_out/r2_l2fetch_default.err
# index build: launch list + full capture of the build kernels (first build only: -c bounds it)
ncu --set full --clock-control none --import-source on -k regex:"^(k_idx_pass|k_idx_sort_buckets|k_idx_dirx|k_idx_split_y)" -c 5 -f -o gpurun_out/r2_idx_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_idx.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^k_feat_genome" -s 3 -c 1 -f -o gpurun_out/r2_featg_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_featg.log 2>&1
tail -3 gpurun_out/r2_ncu_idx.log | cut -c1-200 >> gpurun_out/r2_tests1.log
cat gpurun_out/r2_tests1.log
