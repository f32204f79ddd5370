cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/g40_bench.json 2> gpurun_out/g40_bench.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g40_bench_ref.json 2> gpurun_out/g40_bench_ref.err; echo ref rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/g40_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['copies_only_ceiling_value'], 'parity', d['parity_mode']['ms_per_step'], d['clocks'])
r=json.load(open('gpurun_out/g40_bench_ref.json')); print('ref', r['value'], 'cpu_baseline', d['cpu_baseline']['value'])
PY
TFM_EV_CASES="cubic cubic_shift cubic_ldg jacobian_polar jacobian_fd jacobian_hanning" bash tools/_evidence.sh 2>&1 | tail -8
