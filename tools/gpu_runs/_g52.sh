cd $GRAFT_REPO_ROOT
for v in "" transform_b200/lib/variants/libtfm_b200_k2p_mb5.so; do
echo "=== lib ${v:-default}"
TFM_B200_LIB=${v:+$PWD/$v} timeout 600 python tools/check_k2.py 0 2>&1 | tail -4
done
