python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for g in 1 4 8 32; do echo "group=$g"; LNR_EXTEND_GROUP=$g python bench.py --steps 4 --warmup 2 --no-cpu-baseline --streams 1 2>gpurun_out/err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), {k:round(v['ms_per_launch'],2) for k,v in d['kernels'].items() if k.startswith('k_map')})"; tail -2 gpurun_out/err.log; done
