cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/g49_bench4.json 2> gpurun_out/g49.err; echo rc=$?; tail -2 gpurun_out/g49.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/g49_bench4.json'))
s=d['sharded']
print(d['n_gpus'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), round(d['e2e']['copies_only_ceiling_value']), d['clocks'])
print('sharded lin', s['linear']['ms'], s['linear']['equals_unsharded'], 'jac', s['jacobian']['ms'], s['jacobian']['equals_unsharded'], s['config4_batch'])
PY
