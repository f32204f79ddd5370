cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for extra in "--no-extra" ""; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 $extra > gpurun_out/g42_bench8_${extra:+noextra}.json 2> gpurun_out/g42.err; echo rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/g42_bench8_${extra:+noextra}.json'))
s=d['sharded']
print("extra='$extra'", round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['copies_only_ceiling_value']), 'sharded lin', round(s['linear']['ms'],3), 'jac', round(s['jacobian']['ms'],3), 'cfg4', round(s['config4_batch']['ms'],3))
PY
done
