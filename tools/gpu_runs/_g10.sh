set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/bench_k1_maps.py > gpurun_out/r2j_k1maps.log 2>&1; cat gpurun_out/r2j_k1maps.log
timeout 900 python -m pytest tests/test_gpu_headline.py -q -m gpu -x -k "linear" 2>&1 | tail -15
