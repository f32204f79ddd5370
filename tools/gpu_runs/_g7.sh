set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/prof_k2.py polar > gpurun_out/r2g_plain_polar.log 2>&1 && {
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_jac -s 2 -c 1 -o /tmp/k2_polar python tools/prof_k2.py polar > gpurun_out/r2g_ncu_polar.log 2>&1
  ncu -i /tmp/k2_polar.ncu-rep --page raw --csv > gpurun_out/r2g_k2_polar_raw.csv 2>/dev/null
  ncu -i /tmp/k2_polar.ncu-rep --page source --csv > gpurun_out/r2g_k2_polar_src.csv 2>/dev/null
}
