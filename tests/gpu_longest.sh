LNR_BENCH_BATCH=32768 LNR_LONGEST_PROFILE=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams 1 2>gpurun_out/err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_cycles_last_batch']; names=list(s.keys()); v=list(s.values()); print({k:round(x/1.965e6,3) for k,x in list(s.items())[:12]}); print('total_ms', v[12]/1.965e6, 'read', v[13], 'L', v[14], 'n_raw', v[15])"
tail -2 gpurun_out/err.log
