# round 2, call 52: coalesced k_scan_apply (a warp scans 512 consecutive items, 4 per lane and chunk): parity subset, bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q -x -k "dindex or hindex_build or apxmap_stages or against_unmodified or fuzz or random or records or N_runs" 2>&1 | tail -4 > gpurun_out/r2_tests52.log
cat gpurun_out/r2_tests52.log
LNR_BENCH_NO_SMALL=1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench52.json 2> gpurun_out/r2_bench52.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench52.json'))
k=d['kernels_one_thread']
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), 'scan_seeds', round(k['k_scan_seeds']['ms_per_launch'],3), 'one-thread step', round(d['roofline']['whole_step']['ms_per_step_one_thread'],2), d['roofline']['index']['kernels_ms'].get('k_scan_dir'), d['roofline']['index']['seconds'])
PY
