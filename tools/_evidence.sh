# Evidence run on the GPU box: one ncu --set full capture per kernel case, raw pages exported;
# plus the launch list of the bench command.  Summarised here by tools/ncu_evidence.py.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
declare -A RX=( [linear_shift]=k_resample [linear_named]=k_resample [linear_rot45]=k_resample [nearest]=k_resample [cubic]=k_resample [cubic_shift]=k_resample [cubic_ldg]=k_resample
  [hann]=k_resample [lanczos]=k_resample [linear_polar]=k_resample [linear_parity]=k_resample
  [jacobian_polar]=k_jac [jacobian_polar_rows]=k_jac [jacobian_polar_single]=k_jac [jacobian_constj]=k_jac [jacobian_fd]=k_jac [jacobian_parity]=k_jac [jacobian_hanning]=k_jac
  [wcs_batch]=k_resample [wcs_single]=k_resample [apply_affine]=k_apply [apply_radial]=k_apply [apply_nd]=k_apply [fits_m32]=k_fits [fits_i16]=k_fits )
CASES=${TFM_EV_CASES:-"linear_shift linear_named linear_rot45 nearest cubic cubic_shift cubic_ldg hann lanczos linear_polar linear_parity jacobian_polar jacobian_polar_rows jacobian_polar_single jacobian_constj jacobian_fd jacobian_parity jacobian_hanning wcs_batch wcs_single apply_affine apply_radial apply_nd fits_m32 fits_i16"}
for c in $CASES; do
  skip=3; [ "$c" = jacobian_parity ] && skip=1
  timeout 300 python tools/prof_all.py $c > gpurun_out/ev_plain_$c.log 2>&1 || { echo "plain run failed: $c"; tail -3 gpurun_out/ev_plain_$c.log; continue; }
  timeout 600 ncu --set full --clock-control none -k regex:${RX[$c]} -s $skip -c 1 -o /tmp/ev_$c python tools/prof_all.py $c > gpurun_out/ev_ncu_$c.log 2>&1
  ncu -i /tmp/ev_$c.ncu-rep --page raw --csv > gpurun_out/ev_$c.raw.csv 2>/dev/null
  echo "$c: $(tail -1 gpurun_out/ev_plain_$c.log) $(wc -c < gpurun_out/ev_$c.raw.csv) bytes"
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/ev_bench_plain.json 2> gpurun_out/ev_bench_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/ev_bench_ncu.log 2>&1
rm -f gpurun_out/ev_plain_*.log gpurun_out/ev_ncu_*.log
ls gpurun_out | grep ev_ | wc -l
