# round 2, call 7 (2 GPUs): balanced sharded build at 2 ranks, record construction, N=2 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded_build.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_tests7.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_bench7_n2.json 2> gpurun_out/r2_bench7_n2.err; echo "bench n2 rc=$?" >> gpurun_out/r2_tests7.log
cat gpurun_out/r2_tests7.log
