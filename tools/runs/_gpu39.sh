cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide or queued_batches" 2>&1 | tail -4
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 14 --warmup 4 --no-cpu --no-e2e --parity 512 --no-strong "$@" > gpurun_out/r03o_$n.json 2> gpurun_out/r03o_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03o_$n.json"))
    print("$n value %.0f ms/step %.1f us/step %.4f frac %.3f parity %s"%(d["value"],d["ms_per_step"],d["us_per_newton_step"],d["roofline"]["frac"],(d["parity"] or {}).get("ok")))
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03o_$n.err | tail -3
}
run base
run wide --mopts wide=1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum"
for v in "prefetch=0" "wide=1"; do
timeout 600 ncu --metrics $M --clock-control none -k regex:tcm_solve_kernel -s 3 -c 1 --csv --log-file gpurun_out/r03o_ncu.csv python tools/profile_solve.py --batch 56832 --max-iter 8 --mopts $v > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r03o_ncu.csv")) if len(r)>10]
h=rows[0]
print("  $v", " ".join("%s=%s"%(dict(zip(h,r))["Metric Name"], dict(zip(h,r))["Metric Value"]) for r in rows[1:]))
PY
done
