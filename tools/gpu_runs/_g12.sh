set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_resample.py tests/test_gpu_fuzz.py tests/test_gpu_remap.py tests/test_gpu_sharding.py -q -m gpu 2>&1 | tail -8
timeout 300 python tools/prof_case.py shift linear > gpurun_out/r2l_plain.log 2>&1 && {
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resample -s 3 -c 1 -o /tmp/k1_pipe python tools/prof_case.py shift linear > gpurun_out/r2l_ncu.log 2>&1
ncu -i /tmp/k1_pipe.ncu-rep --page raw --csv > gpurun_out/r2l_k1_pipe_raw.csv 2>/dev/null
ncu -i /tmp/k1_pipe.ncu-rep --page source --csv > gpurun_out/r2l_k1_pipe_src.csv 2>/dev/null
}
