cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tcd_solve -c 1 -o gpurun_out/r02r_dense python tools/dense_probe.py --batch 296 --repeat 1 > gpurun_out/ncu_dense.log 2>&1
tail -2 gpurun_out/ncu_dense.log
