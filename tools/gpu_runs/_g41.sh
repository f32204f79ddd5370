cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/g41_bench2.json 2> gpurun_out/g41_bench2.err; echo rc=$?; tail -2 gpurun_out/g41_bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/g41_ref2.json 2> gpurun_out/g41_ref2.err; echo ref rc=$?; cut -c1-200 gpurun_out/g41_ref2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/g41_bench2.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['copies_only_ceiling_value'])
s=d['sharded']
for k in ('linear','jacobian'): print(k, round(s[k]['ms'],3), round(s[k]['efficiency'],3), s[k]['equals_unsharded'])
print(s['config4_batch'])
PY
