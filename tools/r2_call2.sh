# round 2, call 2: new seeding (match lists) + packed reads: tests, bench, arena sizes, ncu of the seeding kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_cli_apf.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_tests2.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?" >> gpurun_out/r2_tests2.log
for kb in 2048 1024; do
  LNR_ARENA_KB=$kb LNR_BENCH_BATCH=32768 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/r2_arena_$kb.json 2> gpurun_out/r2_arena_$kb.err
done
LNR_BENCH_BATCH=32768 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/r2_b32k.json 2> gpurun_out/r2_b32k.err
ncu --set full --clock-control none --import-source on -k regex:"^(k_seed_count|k_seed_fill)$" -s 2 -c 2 -f -o gpurun_out/r2_seed_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/r2_ncu_seed.log 2>&1
tail -2 gpurun_out/r2_ncu_seed.log | cut -c1-200 >> gpurun_out/r2_tests2.log
cat gpurun_out/r2_tests2.log
