set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/dbg_k2.py > gpurun_out/r2e_dbg.log 2>&1; cat gpurun_out/r2e_dbg.log
V=transform_b200/lib/variants
timeout 600 python tools/ab_variants.py k2 tunings=0,32768,4096 default=default mb3=$V/libtfm_b200_mb3.so mb2=$V/libtfm_b200_mb2.so tp4=$V/libtfm_b200_tp4.so > gpurun_out/r2e_k2.log 2>&1; cat gpurun_out/r2e_k2.log
timeout 600 python -m pytest tests/test_gpu_jacobian.py -q -m gpu 2>&1 | tail -30 > gpurun_out/r2e_tests.log; tail -5 gpurun_out/r2e_tests.log
