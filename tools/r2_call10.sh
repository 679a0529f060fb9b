# round 2, call 10 (8 GPUs): sharded build at 4 ranks (tests), N = 8 and N = 4 bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded_build.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_tests10.log
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/r2_bench10_n$n.json 2> gpurun_out/r2_bench10_n$n.err; echo "bench n$n rc=$?" >> gpurun_out/r2_tests10.log
done
LNR_BENCH_NO_NUMA=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r2_bench10_n8_nonuma.json 2> gpurun_out/r2_bench10_n8_nonuma.err; echo "bench n8 (no NUMA binding) rc=$?" >> gpurun_out/r2_tests10.log
nproc >> gpurun_out/r2_tests10.log; (numactl -H 2>/dev/null || lscpu | grep -i numa) >> gpurun_out/r2_tests10.log
for g in 0 1 2 3 4 5 6 7; do cat /sys/bus/pci/devices/$(nvidia-smi -i $g --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//')/numa_node 2>/dev/null; done | tr '\n' ' ' >> gpurun_out/r2_tests10.log
cat gpurun_out/r2_tests10.log
