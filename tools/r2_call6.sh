# round 2, call 6 (2 GPUs): sharded build inside the C ABI at 2 ranks, hybrid drop-in test, N=1 and N=2 bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded_build.py tests/test_gpu_hybrid.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_tests6.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench6_n1.json 2> gpurun_out/r2_bench6_n1.err; echo "bench n1 rc=$?" >> gpurun_out/r2_tests6.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_bench6_n2.json 2> gpurun_out/r2_bench6_n2.err; echo "bench n2 rc=$?" >> gpurun_out/r2_tests6.log
cat gpurun_out/r2_tests6.log
