cd $GRAFT_REPO_ROOT
timeout 600 python tools/ab_variants.py k2 tunings=0 default=default 2>&1 | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_jacobian.py tests/test_gpu_sharding.py tests/test_gpu_fuzz.py -q -m gpu -x 2>&1 | tail -3
