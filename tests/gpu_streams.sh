for s in 1 2 3; do echo "streams=$s"; python bench.py --steps 6 --warmup 2 --no-cpu-baseline --streams $s 2>gpurun_out/err_$s.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), round(d['kernels']['k_map_hits']['ms_per_launch'],2))"; tail -2 gpurun_out/err_$s.log; done
