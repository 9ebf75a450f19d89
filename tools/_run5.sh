P="python tools/profile_solve.py"
for ls in 8 4 2; do
$P --max-iter 0 --batch 65536 --lockstep $ls 2>&1 | tail -1
$P --max-iter 0 --batch 4736 --lockstep $ls 2>&1 | tail -1
done
$P --max-iter 0 --batch 65536 --lockstep 8 --max-loads 44 --warps-per-sm 16 2>&1 | tail -1
SEC="--section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SpeedOfLight --section InstructionStats --section LaunchStats --section Occupancy"
timeout 400 ncu $SEC --clock-control none -k regex:tcm_solve -c 1 -o /tmp/rep $P --batch 37888 --max-iter 2 --lockstep 8 2>&1 | tail -1
ncu -i /tmp/rep.ncu-rep --page raw --csv > gpurun_out/r01e_ncu_quad_prefetch_lockstep8.csv
timeout 400 ncu $SEC --clock-control none -k regex:tcm_solve -c 1 -o /tmp/rep1 $P --batch 4736 --max-iter 2 2>&1 | tail -1
ncu -i /tmp/rep1.ncu-rep --page raw --csv > gpurun_out/r01e_ncu_quad_prefetch_lonewarp.csv
