cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S=$(date +%s)
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/g58_bench.json 2> gpurun_out/g58.err; echo rc=$? t=$(( $(date +%s) - S ))s; tail -2 gpurun_out/g58.err
python -c "
import json; d=json.load(open('gpurun_out/g58_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['linear_kernel']['frac'], d['cpu_baseline']['value'], d['clocks'])"
