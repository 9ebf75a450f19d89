cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 100 python tools/lone_warp_latency.py --loaded 2>&1 | grep -v attached | tail -5
