cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./tools/bench/pfgran > gpurun_out/r03b_pfgran.txt 2>&1; cat gpurun_out/r03b_pfgran.txt
run() { # name, mopts
  n=$1; shift
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --parity 1024 --no-strong "$@" > gpurun_out/r03b_$n.json 2> gpurun_out/r03b_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03b_$n.json"))
    print("$n value %.0f ms/step %.1f frac %.3f parity %s issued %.3e"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],(d["parity"] or {}).get("ok"),d["roofline"]["issued_bytes_per_launch"]))
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03b_$n.err | tail -3
}
run base
run ls12 --mopts lockstep=12
run ls6 --mopts lockstep=6
run be2 --mopts barrier_every=2
run be4 --mopts barrier_every=4
run pf0 --mopts prefetch=0
run sc75pf0 --mopts scache=75,prefetch=0
run sc75 --mopts scache=75
