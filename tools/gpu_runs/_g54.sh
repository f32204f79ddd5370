cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/g54_bench.json 2> gpurun_out/g54.err; echo rc=$?; tail -2 gpurun_out/g54.err
python -c "
import json; d=json.load(open('gpurun_out/g54_bench.json')); print(d['value'], d['roofline']['ncu'], d['roofline']['linear_kernel']['ncu'], d['cpu_baseline']['value'])"
