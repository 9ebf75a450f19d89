cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 120 ./tools/bench/ifetch > gpurun_out/r04e_ifetch.txt 2>&1; cat gpurun_out/r04e_ifetch.txt
