cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" 1; do
TFM_BENCH_NO_NUMA=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/r2p_bench_$v.json 2>/dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/r2p_bench_$v.json'))
print("no_numa='$v'", "value", round(d['value']), "ms", round(d['ms_per_step'],4), "e2e", round(d['e2e']['value']), "sync", round(d['e2e']['synchronous_calls_value']), d['e2e']['host_affinity'], "parity e2e", round(d['parity_mode']['e2e_value']))
PY
done
