P="python tools/profile_solve.py"
timeout 120 tools/bench/membench2 > gpurun_out/r01d_membench2.txt; cat gpurun_out/r01d_membench2.txt
$P --max-iter 0 --batch 65536 --repeat 2 2>&1 | tail -1
$P --max-iter 0 --batch 4736 2>&1 | tail -1
$P --max-iter 0 --batch 65536 --repeat 2 --max-loads 44 --warps-per-sm 16 2>&1 | tail -1
$P --max-iter 0 --batch 4736 --max-loads 44 --warps-per-sm 16 2>&1 | tail -1
SEC="--section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SpeedOfLight --section InstructionStats --section LaunchStats --section Occupancy"
timeout 400 ncu $SEC --clock-control none -k regex:tcm_solve -c 1 -o /tmp/rep $P --batch 37888 --max-iter 2 2>&1 | tail -1
ncu -i /tmp/rep.ncu-rep --page raw --csv > gpurun_out/r01d_ncu_quad_prefetch.csv
timeout 400 ncu $SEC --clock-control none -k regex:tcm_solve -c 1 -o /tmp/rep16 $P --batch 75776 --max-iter 2 --max-loads 44 --warps-per-sm 16 2>&1 | tail -1
ncu -i /tmp/rep16.ncu-rep --page raw --csv > gpurun_out/r01d_ncu_quad_prefetch_w16.csv
