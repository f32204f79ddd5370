set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra > gpurun_out/r2m_bench2.json 2> gpurun_out/r2m_bench2.err; echo rc=$?; tail -5 gpurun_out/r2m_bench2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_bench2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(json.dumps(d.get('sharded'), indent=1))
PY
