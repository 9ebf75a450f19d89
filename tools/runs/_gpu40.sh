cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"
prof() { # batch mopts
timeout 600 ncu --metrics $M --clock-control none -k regex:tcm_solve_kernel -s 3 -c 1 --csv --log-file gpurun_out/r03p_ncu.csv python tools/profile_solve.py --batch $1 --max-iter 8 --mopts $2 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r03p_ncu.csv")) if len(r)>10]
h=rows[0]
d={dict(zip(h,r))["Metric Name"]: float(dict(zip(h,r))["Metric Value"].replace(",","")) for r in rows[1:]}
t=d["gpu__time_duration.sum"]; by=d["dram__bytes_read.sum"]+d["dram__bytes_write.sum"]
print("  $2 B=$1: %.2f ms, %.3f us per instance-phase, DRAM %.1f GB (%.2f MB/inst) %.2f TB/s, inst %.3g, stalls lsb %.1f noinst %.1f ssb %.1f mio %.1f"%(t/1e6, t/1e3/$1, by/1e9, by/1e6/$1, by/t/1e3, d["smsp__inst_executed.sum"], d["smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"], d["smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"], d["smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"], d["smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]))
PY
}
prof 56832 prefetch=0
prof 113664 pair=1,max_loads=36,prefetch=0
prof 113664 pair=1,max_loads=36,scache=37
prof 113664 pair=1,max_loads=30,scache=37
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 14 --warmup 4 --no-cpu --no-e2e --parity 512 --no-strong "$@" > gpurun_out/r03p_$n.json 2> gpurun_out/r03p_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03p_$n.json"))
    print("$n value %.0f ms/step %.1f us/step %.4f frac %.3f parity %s"%(d["value"],d["ms_per_step"],d["us_per_newton_step"],d["roofline"]["frac"],(d["parity"] or {}).get("ok")), d["config"]["pipeline"][:10])
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03p_$n.err | tail -3
}
run pairsc --mopts pair=1,max_loads=36,scache=37
run pairsc30 --mopts pair=1,max_loads=30,scache=37
