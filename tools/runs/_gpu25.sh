cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, mopts
  n=$1; shift
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --parity 1024 --no-strong "$@" > gpurun_out/r03a_$n.json 2> gpurun_out/r03a_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03a_$n.json"))
    print("$n value %.0f ms/step %.1f frac %.3f parity %s issued %.3e"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],(d["parity"] or {}).get("ok"),d["roofline"]["issued_bytes_per_launch"]))
except Exception as e:
    print("$n FAILED", e)
PY
  tail -3 gpurun_out/r03a_$n.err
}
run base
run sc75 --mopts scache=75
run sc113w8ml96 --mopts scache=113,warps_per_sm=8,max_loads=96
run sc113w8ml72 --mopts scache=113,warps_per_sm=8,max_loads=72
run sc56w16 --mopts scache=56,warps_per_sm=16,max_loads=52
run sc75ml64 --mopts scache=75,max_loads=64
