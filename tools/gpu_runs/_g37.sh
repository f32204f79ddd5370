cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/g37_bench.json 2> gpurun_out/g37_bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/g37_bench.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['copies_only_ceiling_value'], 'parity', d['parity_mode']['ms_per_step'])
b=d['breakdown']
for k in ('linear_affine_fp32_ms','jacobian_radial_fp32_ms','config2_4096_linear_fp32_ms','config2_4096_linear_fp32_roofline_frac','config2_4096_nearest_fp32_roofline_frac','config4_64x4096_tan2car_linear_fp32_roofline_frac','jacobian_radial_fp32_finite_difference_kernel_ms','linear_shift_8192_fp32_roofline_frac'): print(k, b.get(k))
PY
