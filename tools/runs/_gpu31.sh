cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
M="gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"
prof() { # name batch skip mopts
  n=$1; b=$2; sk=$3; shift; shift; shift
  timeout 600 ncu --metrics $M --clock-control none -k regex:tcm_solve_kernel -s $sk -c 1 --csv --log-file gpurun_out/r03g_ncu_$n.csv python tools/profile_solve.py --batch $b --max-iter 8 "$@" > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r03g_ncu_$n.csv")) if len(r)>10]
h=rows[0]; 
print("  $n", " ".join("%s=%s"%(dict(zip(h,r))["Metric Name"].replace("smsp__average_warps_issue_stalled_","stall_").replace("_per_issue_active.ratio","").replace(".sum",""), dict(zip(h,r))["Metric Value"]) for r in rows[1:]))
PY
}
run() { # name, args
  n=$1; shift
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --parity 1024 --no-strong "$@" > gpurun_out/r03g_$n.json 2> gpurun_out/r03g_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03g_$n.json"))
    print("$n value %.0f ms/step %.1f frac %.3f parity %s issued %.3e"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],(d["parity"] or {}).get("ok"),d["roofline"]["issued_bytes_per_launch"]))
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03g_$n.err | tail -3
}
prof pf0 56832 3 --mopts prefetch=0
prof pf0sl 56832 3 --mopts prefetch=0,sort_loads=1
run pf0 --mopts prefetch=0
run pf0sl --mopts prefetch=0,sort_loads=1
