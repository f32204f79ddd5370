cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 100 python bench.py --steps 20 --warmup 3 > gpurun_out/g62_bench.json 2> gpurun_out/g62.err; echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/g62_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('copies_only_ceiling_value'), d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])"
