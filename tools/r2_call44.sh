# round 2, call 44: the whole GPU suite and the default bench line of the final build
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/r2_tests44.log 2>&1
cat gpurun_out/r2_tests44.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench44.json 2> gpurun_out/r2_bench44.err ) 2> gpurun_out/r2_bench44.time; echo "bench rc=$?"; cat gpurun_out/r2_bench44.time
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench44.json'))
print(round(d['value']), round(d['e2e']['value']), round(d['e2e_dna5']['value']), round(d['ms_per_step'],2), d['parity'], d['roofline']['kernel'], round(d['roofline']['frac'],3), round(d['roofline']['whole_step']['frac'],3), d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'])
PY
