LNR_BENCH_BATCH=32768 LNR_TRACE=1 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --streams 2 2>&1 >/dev/null | grep "lnr trace" | tail -8
