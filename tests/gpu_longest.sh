LNR_LONGEST_PROFILE=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 2>gpurun_out/err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_cycles_last_batch']; print({k:round(v/1.965e6,3) for k,v in s.items()})"
tail -2 gpurun_out/err.log
