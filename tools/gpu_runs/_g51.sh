cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/g51_bench.json 2> gpurun_out/g51_bench.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g51_bench_ref.json 2> gpurun_out/g51_bench_ref.err; echo ref rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/g51_bench.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['copies_only_ceiling_value'], d['clocks'])
print('cpu_baseline', d['cpu_baseline']['value'], d['cpu_baseline']['kind']['linear'][:60])
r=json.load(open('gpurun_out/g51_bench_ref.json')); print('ref', r['value'], r['cpu_baseline']['linear_mpix_s'], r['cpu_baseline']['jacobian_mpix_s'])
PY
