cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/g30_bench.json 2> gpurun_out/g30_bench.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g30_bench_ref.json 2> gpurun_out/g30_bench_ref.err; echo ref rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/g30_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['kernels'], 'e2e', d['e2e']['value'], 'parity', d['parity_mode']['ms_per_step'], d['parity_mode']['e2e_value'])
b=d['breakdown']
for k in ('jacobian_radial_fp32_ms','jacobian_downscale_8192_fp32_ms','config4_64x4096_tan2car_linear_fp32_ms','config4_64x4096_tan2car_linear_fp32_roofline_frac','jacobian_radial_fp64_parity_ms','linear_shift_8192_fp32_roofline_frac'): print(k, b.get(k))
r=json.load(open('gpurun_out/g30_bench_ref.json'))
print('ref', r['value'], r['cpu_baseline']['cores'], r['ms_per_step'])
print('cpu_baseline', d['cpu_baseline']['value'])
PY
