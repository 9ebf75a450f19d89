P="python tools/profile_solve.py"
python -m pytest tests -m gpu -x -q -s -k compiled 2>&1 | grep -v "^init\|^/root/ref" | tail -7
$P --max-iter 0 --batch 65536 --repeat 2 2>&1 | tail -1
$P --max-iter 0 --batch 4736 2>&1 | tail -1
$P --max-iter 0 --batch 262144 --workload mpc_dcmotor --repeat 2 2>&1 | tail -1
$P --max-iter 0 --batch 65536 --workload l1l2 --repeat 2 2>&1 | tail -1
SEC="--section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SpeedOfLight --section InstructionStats --section LaunchStats --section Occupancy"
timeout 400 ncu $SEC --clock-control none -k regex:tcm_solve -c 1 -o /tmp/rep $P --batch 37888 --max-iter 2 2>&1 | tail -1
ncu -i /tmp/rep.ncu-rep --page raw --csv > gpurun_out/r01f_ncu_quad_merge8_div.csv
