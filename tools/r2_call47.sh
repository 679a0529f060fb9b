# round 2, call 47: side lines of the final build: HiFi-like reads (15 kb, 1 % error), HIndex seeding (-i 2) on a 50-Mbase genome
mkdir -p gpurun_out
export LNR_BENCH_NO_SMALL=1
LNR_BENCH_PROFILE=hifi timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench47_hifi_side.json 2> gpurun_out/r2_bench47_hifi.err; echo "hifi rc=$?"
LNR_BENCH_INDEX=2 LNR_BENCH_GENOME=50000000 timeout 200 python bench.py --steps 4 --warmup 3 --batch-reads 16384 --no-cpu-baseline > gpurun_out/r2_bench47_hindex_50m_side.json 2> gpurun_out/r2_bench47_hindex.err; echo "hindex rc=$?"
python - <<PY
import json
for f in ('hifi_side','hindex_50m_side'):
    d=json.load(open('gpurun_out/r2_bench47_%s.json' % f))
    k=d['kernels_one_thread']
    print(f, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), {x: round(v['ms_per_launch'],2) for x,v in k.items() if v['ms_per_launch']>0.5})
PY
