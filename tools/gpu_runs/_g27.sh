cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-extra > gpurun_out/g27_bench2.json 2> gpurun_out/g27_bench2.err; echo rc=$?; tail -3 gpurun_out/g27_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/g27_bench2.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print(json.dumps(d.get('sharded'), indent=1)[:3000])
PY
