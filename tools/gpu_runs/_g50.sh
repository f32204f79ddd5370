cd $GRAFT_REPO_ROOT
ls /root/reference 2>&1 | head -2
python - <<'PY'
import sys
sys.path.insert(0,'oracle')
import ref_import
ref = ref_import.load_reference_compiled()
print("compiled reference loaded:", ref.Composition, ref.interpND)
PY
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['cpu_baseline']['kind']['linear'][:120], d['cpu_baseline']['linear_mpix_s'], d['cpu_baseline']['jacobian_mpix_s'])"
