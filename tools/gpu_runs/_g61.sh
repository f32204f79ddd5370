cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/g61_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python bench.py --steps 20 --warmup 3 > gpurun_out/g61_bench.json 2> gpurun_out/g61.err; echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/g61_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['achieved'], d['clocks'])"
timeout 60 python tools/prof_all.py jacobian_polar > gpurun_out/ev_plain_jp.log 2>&1 && \
timeout 150 ncu --set full --clock-control none -k regex:k_jac -s 3 -c 1 -o /tmp/ev_jp python tools/prof_all.py jacobian_polar > gpurun_out/ev_ncu_jp.log 2>&1
ncu -i /tmp/ev_jp.ncu-rep --page raw --csv > gpurun_out/ev_jacobian_polar.raw.csv 2>/dev/null
tail -1 gpurun_out/ev_plain_jp.log; wc -c gpurun_out/ev_jacobian_polar.raw.csv
