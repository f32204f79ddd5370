set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_headline.py tests/test_gpu_jacobian.py tests/test_gpu_resample.py tests/test_gpu_sharding.py tests/test_gpu_golden.py -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2b_tests.log
cat gpurun_out/r2b_tests.log
timeout 300 python tools/bench_k1_maps.py > gpurun_out/r2b_k1maps.log 2>&1; cat gpurun_out/r2b_k1maps.log
timeout 300 python tools/ab_variants.py k2 default=default > gpurun_out/r2b_k2.log 2>&1; cat gpurun_out/r2b_k2.log
for c in polar affine; do
  timeout 300 python tools/prof_k2.py $c > gpurun_out/r2b_plain_$c.log 2>&1 || continue
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_jac -s 2 -c 1 -o /tmp/k2_$c python tools/prof_k2.py $c > gpurun_out/r2b_ncu_$c.log 2>&1
  ncu -i /tmp/k2_$c.ncu-rep --page raw --csv > gpurun_out/r2b_k2_${c}_raw.csv 2>/dev/null
  ncu -i /tmp/k2_$c.ncu-rep --page source --csv > gpurun_out/r2b_k2_${c}_src.csv 2>/dev/null
done
timeout 300 python tools/prof_case.py shift > gpurun_out/r2b_plain_shift.log 2>&1 && {
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resample -s 3 -c 1 -o /tmp/k1_shift python tools/prof_case.py shift > gpurun_out/r2b_ncu_shift.log 2>&1
ncu -i /tmp/k1_shift.ncu-rep --page raw --csv > gpurun_out/r2b_k1_shift_raw.csv 2>/dev/null
ncu -i /tmp/k1_shift.ncu-rep --page source --csv > gpurun_out/r2b_k1_shift_src.csv 2>/dev/null
}
ls -la gpurun_out | tail -12
