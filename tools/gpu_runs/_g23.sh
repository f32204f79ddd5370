cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/g23_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/g23_bench.json 2> gpurun_out/g23_bench.err
tail -3 gpurun_out/g23_tests.log; cut -c1-1500 gpurun_out/g23_bench.json
