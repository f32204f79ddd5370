cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for numa in "" 1; do
TFM_BENCH_NO_NUMA=$numa timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/host_ceiling.py 2>/dev/null | tail -1
done
TFM_BENCH_NO_NUMA= timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 tools/host_ceiling.py 2>/dev/null | tail -1
TFM_BENCH_NO_NUMA= timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/host_ceiling.py 2>/dev/null | tail -1
timeout 120 python tools/host_ceiling.py 2>/dev/null | tail -1
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" | head -8
