cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/g26_bench.json 2> gpurun_out/g26_bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/g26_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['kernels'], d['e2e']['value'], d['breakdown']['jacobian_radial_fp32_ms'], d['breakdown']['jacobian_downscale_8192_fp32_ms'])
PY
