cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_sharding.py tests/test_gpu_jacobian.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/sharded_nccl_demo.py 2>&1 | grep -v "^\*\|OMP_NUM" | grep "polar\|NCCL" | tee gpurun_out/r2n_sharded2.log
