cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 150 python tools/lone_warp_latency.py 2>&1 | grep -v attached | tail -6
