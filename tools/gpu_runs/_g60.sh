cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=transform_b200/lib/variants
timeout 200 python tools/ab_variants.py k2 default=default db4_mb3=$V/libtfm_b200_db4_mb3.so db5_mb3=$V/libtfm_b200_db5_mb3.so db6_mb3=$V/libtfm_b200_db6_mb3.so db4_mb3_again=$V/libtfm_b200_db4_mb3.so 2>&1 | tail -8
