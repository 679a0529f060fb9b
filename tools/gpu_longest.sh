LNR_BENCH_BATCH=${LNR_BENCH_BATCH:-32768} LNR_LONGEST_PROFILE=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 2>gpurun_out/err.log >/dev/null
grep "lnr tail" gpurun_out/err.log | tail -21
