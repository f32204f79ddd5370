cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_sip.py -q -m gpu 2>&1 | tail -40 | tee gpurun_out/g57_sip.log
timeout 300 python -m pytest tests -q -m gpu --deselect tests/test_gpu_sip.py 2>&1 | tail -15 | tee gpurun_out/g57_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
