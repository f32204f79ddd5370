cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/g32_bench8.json 2> gpurun_out/g32_bench8.err; echo rc=$?; tail -3 gpurun_out/g32_bench8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/g32_bench8.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('synchronous_calls_value'), d['clocks'])
s=d.get('sharded') or {}
for k in ('linear','jacobian'):
    print(k, {kk: s[k][kk] for kk in ('ms','single_gpu_ms','speedup','efficiency','split','link_bound_ms','equals_unsharded')})
print(s.get('config4_batch'))
PY
