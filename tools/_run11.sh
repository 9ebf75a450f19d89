P="python tools/profile_solve.py --workload mpc_dcmotor --batch 262144 --max-iter 0 --repeat 2"
$P 2>&1 | tail -1 | cut -c1-120
$P --max-loads 44 --warps-per-sm 16 2>&1 | tail -1 | cut -c1-120
$P --max-loads 64 --warps-per-sm 12 2>&1 | tail -1 | cut -c1-120
$P --max-loads 28 --warps-per-sm 24 2>&1 | tail -1 | cut -c1-120
SEC="--section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SpeedOfLight --section InstructionStats --section LaunchStats --section Occupancy"
timeout 300 ncu $SEC --clock-control none -k regex:tcm_solve -c 1 -o /tmp/repd python tools/profile_solve.py --workload mpc_dcmotor --batch 262144 --max-iter 0 2>&1 | tail -1 | cut -c1-100
ncu -i /tmp/repd.ncu-rep --page raw --csv > gpurun_out/r01j_ncu_dcmotor.csv
