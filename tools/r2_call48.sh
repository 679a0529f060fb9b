# round 2, call 48: binning filter picks the dense scheme when most anchors are candidates (counted from the sketch):
# parity subset (DIndex + HIndex + fuzz), HIndex side line, headline bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q -x -k "apxmap_stages or against_unmodified or hindex_apxmap or fuzz or random or scratch_overflow or one_kernel" 2>&1 | tail -4 > gpurun_out/r2_tests48.log
cat gpurun_out/r2_tests48.log
export LNR_BENCH_NO_SMALL=1
LNR_BENCH_INDEX=2 LNR_BENCH_GENOME=50000000 timeout 200 python bench.py --steps 4 --warmup 3 --batch-reads 16384 --no-cpu-baseline > gpurun_out/r2_bench48_hindex_50m_side.json 2> gpurun_out/r2_bench48_hindex.err; echo "hindex rc=$?"
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench48.json 2> gpurun_out/r2_bench48.err; echo "bench rc=$?"
python - <<PY
import json
for f in ('_hindex_50m_side',''):
    d=json.load(open('gpurun_out/r2_bench48%s.json' % f))
    k=d['kernels_one_thread']
    print(f, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), 'sort', round(k['k_hits_sort']['ms_per_launch'],2))
PY
