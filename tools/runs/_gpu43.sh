cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_misc_per_issue_active.ratio,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction_lookup_miss.sum,lts__t_requests_srcunit_gcc.sum,lts__t_requests_srcunit_gcc_lookup_miss.sum,sm__cycles_active.max"
for b in 32 128; do
timeout 600 ncu --metrics $M --clock-control none -k regex:tcm_solve_kernel -s 3 -c 1 --csv --log-file gpurun_out/r03t_ncu_b$b.csv python tools/profile_solve.py --batch $b --max-iter 8 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r03t_ncu_b$b.csv")) if len(r)>10]
h=rows[0]
print("batch $b:", " ".join("%s=%s"%(dict(zip(h,r))["Metric Name"].replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio",""), dict(zip(h,r))["Metric Value"]) for r in rows[1:]))
PY
done
python tools/profile_solve.py --batch 32 --max-iter 40 --repeat 2 2>&1 | tail -1 | cut -c1-200
