cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, args
  n=$1; shift
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --parity 512 --no-strong "$@" > gpurun_out/r03h_$n.json 2> gpurun_out/r03h_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03h_$n.json"))
    print("$n value %.0f ms/step %.1f us/step %.4f frac %.3f parity %s"%(d["value"],d["ms_per_step"],d["us_per_newton_step"],d["roofline"]["frac"],(d["parity"] or {}).get("ok")))
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03h_$n.err | tail -3
}
run q56832 --mopts prefetch=0 --batch 56832
run q113664 --mopts prefetch=0 --batch 113664 --pipeline 3
run q65536p8 --mopts prefetch=0 --pipeline 8
for w in l1l2 mpc_dcmotor sls; do
run ${w}_base --workload $w
run ${w}_pf0 --workload $w --mopts prefetch=0
done
