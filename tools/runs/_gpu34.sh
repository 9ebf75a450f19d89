cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 14 --warmup 4 --no-cpu --no-e2e --parity 256 --no-strong --pipeline 7 "$@" > gpurun_out/r03j_$n.json 2> gpurun_out/r03j_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03j_$n.json"))
    print("$n value %.0f ms/step %.1f us/step %.4f frac %.3f parity %s"%(d["value"],d["ms_per_step"],d["us_per_newton_step"],d["roofline"]["frac"],(d["parity"] or {}).get("ok")))
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03j_$n.err | tail -3
}
run ml72 --mopts prefetch=0
run ml64 --mopts prefetch=0,max_loads=64
run ml80 --mopts prefetch=0,max_loads=80
run ml88 --mopts prefetch=0,max_loads=88
run ls2 --mopts prefetch=0,lockstep=2
run ls3 --mopts prefetch=0,lockstep=3
