P="python tools/profile_solve.py"
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^init\|^/root/ref" | tail -12
$P --max-iter 0 --batch 65536 --repeat 2 2>&1 | tail -2
$P --max-iter 0 --batch 4736 --repeat 2 2>&1 | tail -1
$P --max-iter 0 --batch 262144 --repeat 2 --workload mpc_dcmotor 2>&1 | tail -1
SEC="--section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SpeedOfLight --section InstructionStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis"
timeout 400 ncu $SEC --clock-control none -k regex:tcm_solve -c 1 -o /tmp/rep $P --batch 37888 --max-iter 2 2>&1 | tail -2
ncu -i /tmp/rep.ncu-rep --page raw --csv > gpurun_out/r01c_ncu_quad_loadsfirst.csv
