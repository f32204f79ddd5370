cd $GRAFT_REPO_ROOT
timeout 45 python -m pytest tests/test_gpu_sip.py tests/test_gpu_remap.py tests/test_gpu_apply.py -q -m gpu -x 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
