cd $GRAFT_REPO_ROOT
timeout 600 python tools/check_k2.py 2>&1 | tail -22
timeout 600 python -m pytest tests/test_gpu_jacobian.py tests/test_gpu_headline.py tests/test_gpu_sharding.py -q -m gpu -x 2>&1 | tail -3
