cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
M="smsp__inst_executed.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores,sass__inst_executed_global_loads,sass__inst_executed_global_stores,dram__bytes.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active"
prof() { # name batch mopts
  n=$1; b=$2; shift; shift
  timeout 300 python tools/profile_solve.py --batch $b --max-iter 4 --repeat 3 "$@" 2>/dev/null | tail -1 | cut -c1-160
  timeout 600 ncu --metrics $M --clock-control none -k regex:tcm_solve_kernel -s 1 -c 1 --csv --log-file gpurun_out/r03d_ncu_$n.csv python tools/profile_solve.py --batch $b --max-iter 2 "$@" > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r03d_ncu_$n.csv")) if len(r)>10]
h=rows[0]; 
for r in rows[1:]:
    d=dict(zip(h,r)); print("  $n", d["Metric Name"].replace("smsp__average_warps_issue_stalled_","stall_").replace("_per_issue_active.ratio",""), d["Metric Value"])
PY
}
prof pf0 56832 --mopts prefetch=0
prof pair36 113664 --mopts pair=1,max_loads=36,prefetch=0
prof pair24 113664 --mopts pair=1,max_loads=24,prefetch=0
prof pair28 113664 --mopts pair=1,max_loads=28,prefetch=0
prof pair56w8 75776 --mopts pair=1,max_loads=56,warps_per_sm=8,prefetch=0
prof pf0sc75 56832 --mopts prefetch=0,scache=75
