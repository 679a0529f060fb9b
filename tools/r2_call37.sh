# round 2, call 36-37: seeding kernels at 8 CTAs/SM + 32-ary warp search of the first task (k_seed_count) / task handed over in the
# warp record (k_seed_fill): parity subset, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q -x -k "apxmap_stages or against_unmodified or exhausted or N_runs or fuzz or random or packed or hindex or c0" 2>&1 | tail -5 > gpurun_out/r2_tests37.log
cat gpurun_out/r2_tests37.log
export LNR_BENCH_NO_SMALL=1
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench37.json 2> gpurun_out/r2_bench37.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench37.json'))
k=d['kernels_one_thread']
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), 'count', round(k['k_seed_count']['ms_per_launch'],3), 'fill', round(k['k_seed_fill']['ms_per_launch'],3), 'one-thread step', round(d['roofline']['whole_step']['ms_per_step_one_thread'],2))
PY
