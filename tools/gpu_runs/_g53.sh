cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_apply.py -q -m gpu -x 2>&1 | tail -3
