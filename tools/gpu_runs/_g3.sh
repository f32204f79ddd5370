set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/bench_k1_maps.py > gpurun_out/r2c_k1maps.log 2>&1; cat gpurun_out/r2c_k1maps.log
V=transform_b200/lib/variants
timeout 600 python tools/ab_variants.py k2 tunings=0,4096,1024,2048,16384 default=default mb3=$V/libtfm_b200_mb3.so mb2=$V/libtfm_b200_mb2.so > gpurun_out/r2c_k2.log 2>&1; cat gpurun_out/r2c_k2.log
timeout 1500 python -m pytest tests/test_gpu_headline.py tests/test_gpu_jacobian.py tests/test_gpu_resample.py tests/test_gpu_sharding.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py -q -m gpu 2>&1 | tail -60 > gpurun_out/r2c_tests.log
cat gpurun_out/r2c_tests.log
timeout 300 python tools/prof_k2.py polar > gpurun_out/r2c_plain_polar.log 2>&1 && {
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_jac -s 2 -c 1 -o /tmp/k2_polar python tools/prof_k2.py polar > gpurun_out/r2c_ncu_polar.log 2>&1
  ncu -i /tmp/k2_polar.ncu-rep --page raw --csv > gpurun_out/r2c_k2_polar_raw.csv 2>/dev/null
  ncu -i /tmp/k2_polar.ncu-rep --page source --csv > gpurun_out/r2c_k2_polar_src.csv 2>/dev/null
}
timeout 300 python tools/prof_case.py shift > gpurun_out/r2c_plain_shift.log 2>&1 && {
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resample -s 3 -c 1 -o /tmp/k1_shift python tools/prof_case.py shift > gpurun_out/r2c_ncu_shift.log 2>&1
ncu -i /tmp/k1_shift.ncu-rep --page raw --csv > gpurun_out/r2c_k1_shift_raw.csv 2>/dev/null
ncu -i /tmp/k1_shift.ncu-rep --page source --csv > gpurun_out/r2c_k1_shift_src.csv 2>/dev/null
}
