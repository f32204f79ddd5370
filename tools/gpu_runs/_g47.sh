cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/g47_bench.json 2> gpurun_out/g47.err; echo rc=$?; tail -3 gpurun_out/g47.err
python -c "
import json; d=json.load(open('gpurun_out/g47_bench.json')); print(d['value'], d['ms_per_step'], d['clocks'])"
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks'])"
