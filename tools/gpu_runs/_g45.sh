cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_jacobian.py -q -m gpu -x -k "row_pair" 2>&1 | tail -15
