cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, args
  n=$1; shift
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --parity 1024 --no-strong "$@" > gpurun_out/r03c_$n.json 2> gpurun_out/r03c_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03c_$n.json"))
    print("$n value %.0f ms/step %.1f frac %.3f parity %s issued %.3e"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],(d["parity"] or {}).get("ok"),d["roofline"]["issued_bytes_per_launch"]), d["parity"])
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03c_$n.err | tail -3; grep attached gpurun_out/r03c_$n.err | head -1
}
run sls_pair --workload sls --mopts pair=1,max_loads=14
run dcm_pair --workload mpc_dcmotor --mopts pair=1,max_loads=14
run pf0 --mopts prefetch=0
run pf0be4 --mopts prefetch=0,barrier_every=4
run pf0sc75 --mopts prefetch=0,scache=75
run pair36 --mopts pair=1,max_loads=36,prefetch=0
run pair32be4 --mopts pair=1,max_loads=32,prefetch=0,barrier_every=4
run pair56w8 --mopts pair=1,max_loads=56,warps_per_sm=8,prefetch=0
run pair40 --mopts pair=1,max_loads=40,prefetch=0
run pair36pf1 --mopts pair=1,max_loads=36
