set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resample.py -q -m gpu -x 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo rc=$?; tail -3 gpurun_out/r2i_bench.err; head -c 3000 gpurun_out/r2i_bench.json
