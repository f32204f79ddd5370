cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=transform_b200/lib/variants
timeout 200 python tools/ab_variants.py k2 default=default db4_mb4=$V/libtfm_b200_db4_mb4.so db4_mb3=$V/libtfm_b200_db4_mb3.so db3_mb3=$V/libtfm_b200_db3_mb3.so base_mb3=$V/libtfm_b200_base_mb3.so default2=default 2>&1 | tail -8
