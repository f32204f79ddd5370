set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( timeout 300 python tools/prof_k2.py polar > gpurun_out/r2a_plain.log 2>&1 && timeout 300 python tools/prof_k2.py affine >> gpurun_out/r2a_plain.log 2>&1 && timeout 300 python tools/prof_case.py shift >> gpurun_out/r2a_plain.log 2>&1 ) || { echo PLAIN FAILED; tail -20 gpurun_out/r2a_plain.log; exit 1; }
for c in polar affine; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_jacobian -s 2 -c 1 -o /tmp/k2_$c python tools/prof_k2.py $c > gpurun_out/r2a_ncu_$c.log 2>&1
  ncu -i /tmp/k2_$c.ncu-rep --page raw --csv > gpurun_out/r2a_k2_${c}_raw.csv 2>/dev/null
  ncu -i /tmp/k2_$c.ncu-rep --page source --csv > gpurun_out/r2a_k2_${c}_src.csv 2>/dev/null
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resample -s 3 -c 1 -o /tmp/k1_shift python tools/prof_case.py shift > gpurun_out/r2a_ncu_shift.log 2>&1
ncu -i /tmp/k1_shift.ncu-rep --page raw --csv > gpurun_out/r2a_k1_shift_raw.csv 2>/dev/null
ncu -i /tmp/k1_shift.ncu-rep --page source --csv > gpurun_out/r2a_k1_shift_src.csv 2>/dev/null
ls -la gpurun_out | tail -12
