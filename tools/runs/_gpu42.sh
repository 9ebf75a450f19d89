cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 14 --warmup 4 --no-cpu --no-e2e --parity 256 --no-strong "$@" > gpurun_out/r03s_$n.json 2> gpurun_out/r03s_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03s_$n.json"))
    print("$n value %.0f ms/step %.1f us/step %.4f frac %.3f parity %s"%(d["value"],d["ms_per_step"],d["us_per_newton_step"],d["roofline"]["frac"],(d["parity"] or {}).get("ok")), d["config"]["pipeline"][:10])
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03s_$n.err | tail -2; grep attached gpurun_out/r03s_$n.err | head -1 | cut -c60-200
}
run w14ml52 --mopts lockstep=2,warps_per_sm=14,max_loads=52
run w14ml44 --mopts lockstep=2,warps_per_sm=14,max_loads=44
run w14ml52b56832 --mopts lockstep=2,warps_per_sm=14,max_loads=52 --batch 56832
