# round 2, call 26: warp tracebacks + speculative run filter: parity + bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q -x -k "apxmap_stages or against_unmodified or scratch_overflow or one_kernel or fuzz or random or hindex_apxmap or f1_features or c0" 2>&1 | tail -5 > gpurun_out/r2_tests26.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench26.json 2> gpurun_out/r2_bench26.err; echo "bench rc=$?" >> gpurun_out/r2_tests26.log
cat gpurun_out/r2_tests26.log
