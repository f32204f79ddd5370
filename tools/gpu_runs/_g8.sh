set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=transform_b200/lib/variants
timeout 600 python tools/ab_variants.py k2 tunings=0,2048,16384 default=default mb3=$V/libtfm_b200_mb3.so > gpurun_out/r2h_k2.log 2>&1; cat gpurun_out/r2h_k2.log
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_jacobian.py -q -m gpu 2>&1 | tail -5
