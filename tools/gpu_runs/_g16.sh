cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_sharding.py tests/test_gpu_resample.py -q -m gpu -x 2>&1 | tail -12
