# round 2, call 13: 64-byte dirx entries (56 Y keys): full GPU suite + bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_tests13.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; echo "bench rc=$?" >> gpurun_out/r2_tests13.log
cat gpurun_out/r2_tests13.log
