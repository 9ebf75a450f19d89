cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tcd_solve -c 1 -o /tmp/r02z_dense python tools/dense_probe.py --batch 592 --repeat 1 > gpurun_out/ncu_dense.log 2>&1
tail -2 gpurun_out/ncu_dense.log | cut -c1-300
ncu -i /tmp/r02z_dense.ncu-rep --page source --csv > gpurun_out/r02z_dense_source.csv 2>/dev/null
ncu -i /tmp/r02z_dense.ncu-rep --page raw --csv > /tmp/raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/raw.csv | grep -v pcsamp > gpurun_out/r02z_ncu_tcd_solve_kernel_summary.txt
ls -la gpurun_out/r02z*
