cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/dense_probe.py --batch 1024 --repeat 2 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tcd_ -c 2 -o /tmp/r03l_dense python tools/dense_probe.py --batch 592 --repeat 1 > gpurun_out/ncu_dense.log 2>&1
ncu -i /tmp/r03l_dense.ncu-rep --page raw --csv > /tmp/raw_dense.csv 2>/dev/null
python tools/ncu_summary.py /tmp/raw_dense.csv > gpurun_out/r03l_ncu_dense_summary.txt
ncu -i /tmp/r03l_dense.ncu-rep --page details > gpurun_out/r03l_ncu_dense_details.txt 2>/dev/null
ncu -i /tmp/r03l_dense.ncu-rep --page source --csv > gpurun_out/r03l_ncu_dense_source.csv 2>/dev/null
ls -la gpurun_out/r03l*
