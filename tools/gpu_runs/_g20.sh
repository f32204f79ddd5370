cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/ab_variants.py k2 tunings=0,65536,2048,67584,16384,81920 default=default 2>&1 | cut -c1-900
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_jacobian.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python tools/e2e_probe.py 2>&1 | tail -12
