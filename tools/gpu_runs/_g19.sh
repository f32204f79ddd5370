cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for m in single; do
timeout 300 python tools/prof_wcs.py $m > gpurun_out/r2o_plain_$m.log 2>&1 && {
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resample -s 2 -c 1 -o /tmp/wcs_$m python tools/prof_wcs.py $m > gpurun_out/r2o_ncu_$m.log 2>&1
ncu -i /tmp/wcs_$m.ncu-rep --page raw --csv > gpurun_out/r2o_wcs_${m}_raw.csv 2>/dev/null
ncu -i /tmp/wcs_$m.ncu-rep --page source --csv > gpurun_out/r2o_wcs_${m}_src.csv 2>/dev/null
}
done
