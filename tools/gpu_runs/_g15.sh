cd $GRAFT_REPO_ROOT
timeout 600 python tools/dbg_shard.py 2>&1 | tail -30
