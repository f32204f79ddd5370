cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/g43_bench2.json 2> gpurun_out/g43.err; echo rc=$?; tail -2 gpurun_out/g43.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/g43_bench2.json'))
s=d['sharded']
print(round(d['value']), 'sharded lin', s['linear']['ms'], s['linear']['batches_ms'], 'jac', s['jacobian']['ms'], s['jacobian']['batches_ms'])
PY
