# round 2, call 21: does the end-to-end number want more host threads / smaller batches? (side runs, not the headline)
mkdir -p gpurun_out
LNR_BENCH_NO_SMALL=1 timeout 300 python bench.py --steps 12 --warmup 6 --streams 6 --no-cpu-baseline > gpurun_out/r2_bench21_s6.json 2> gpurun_out/r2_bench21_s6.err; echo "s6 rc=$?"
LNR_BENCH_NO_SMALL=1 timeout 300 python bench.py --steps 16 --warmup 8 --streams 8 --batch-reads 32768 --no-cpu-baseline > gpurun_out/r2_bench21_s8_b32k.json 2> gpurun_out/r2_bench21_s8_b32k.err; echo "s8 rc=$?"
LNR_BENCH_NO_SMALL=1 timeout 300 python bench.py --steps 8 --warmup 4 --streams 2 --no-cpu-baseline > gpurun_out/r2_bench21_s2.json 2> gpurun_out/r2_bench21_s2.err; echo "s2 rc=$?"
