for c in 3 4 6; do echo "ctas_per_sm=$c"; LNR_MAP_CTAS_PER_SM=$c python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), round(d['kernels']['k_map_hits']['ms_per_launch'],2))"; done
