P="python tools/profile_solve.py"
for b in 4736 9472 18944 37888; do $P --lockstep 1 --max-iter 4 --batch $b 2>&1 | tail -1; done
for ls in 4 8; do $P --lockstep $ls --max-iter 4 --batch 37888 2>&1 | tail -1; done
for ls in 4 8; do $P --lockstep $ls --max-iter 0 --batch 65536 2>&1 | tail -1; done
SEC="--section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SpeedOfLight --section InstructionStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis"
for ls in 1 8; do
timeout 400 ncu $SEC --clock-control none -k regex:tcm_solve -c 1 -o /tmp/rep$ls $P --lockstep $ls --batch 37888 --max-iter 2 2>&1 | tail -2
ncu -i /tmp/rep$ls.ncu-rep --page raw --csv > gpurun_out/r01b_ncu_quad_lockstep$ls.csv
done
ls -la gpurun_out
