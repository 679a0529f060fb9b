# round 2, call 14: -c 0 path on the GPU (k_map_c0) against oracle / golden / reference
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c0" 2>&1 | tail -30 > gpurun_out/r2_tests14.log
cat gpurun_out/r2_tests14.log
