cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for m in jacobian_fd; do
timeout 300 python tools/prof_all.py $m > gpurun_out/g25_plain_$m.log 2>&1 && {
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_jac -s 3 -c 1 -o /tmp/p_$m python tools/prof_all.py $m > gpurun_out/g25_ncu_$m.log 2>&1
ncu -i /tmp/p_$m.ncu-rep --page source --csv > gpurun_out/g25_${m}_src.csv 2>/dev/null
}
tail -2 gpurun_out/g25_plain_$m.log
done
