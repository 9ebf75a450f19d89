cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-strong "$@" > gpurun_out/r02g_$n.json 2> gpurun_out/r02g_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02g_$n.json"))
    p=d.get("parity") or {}
    print("$n value %.0f ms/step %.1f frac %.3f launches %d parity %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["gpu_launches"],p.get("ok")))
except Exception as e:
    print("$n FAILED", e)
PY
}
run base --pipeline 4 --parity 0
run ls2 --pipeline 4 --parity 0 --mopts lockstep=2
run ls4 --pipeline 4 --parity 1024 --mopts lockstep=4
run ls8 --pipeline 4 --parity 0 --mopts lockstep=8
run ls8w16 --pipeline 4 --parity 0 --mopts lockstep=8,warps_per_sm=16,max_loads=44
run ls4w12 --pipeline 4 --parity 0 --mopts lockstep=4,warps_per_sm=12,max_loads=64
