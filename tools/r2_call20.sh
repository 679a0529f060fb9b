# round 2, call 20: SWAR Y-key match in k_seed_count: quick parity subset + bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "apxmap_stages or against_unmodified or exhausted or N_runs" 2>&1 | tail -5 > gpurun_out/r2_tests20.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err; echo "bench rc=$?" >> gpurun_out/r2_tests20.log
cat gpurun_out/r2_tests20.log
