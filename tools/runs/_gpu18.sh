cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu --no-e2e --no-strong --parity 0 "$@" > gpurun_out/r02v_$n.json 2> gpurun_out/r02v_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02v_$n.json"))
    print("$n value %.0f ms/step %.1f frac %.3f launches %d"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["gpu_launches"]))
except Exception as e:
    print("$n FAILED", e)
PY
}
run base
run pf2 --mopts prefetch=2
run pf3 --mopts prefetch=3
run t2ml48 --mopts lockstep=4,tiles=2,warps_per_sm=12,max_loads=48
run t2ml72 --mopts lockstep=4,tiles=2,warps_per_sm=12,max_loads=72
run t2w8 --mopts lockstep=4,tiles=2,warps_per_sm=8,max_loads=96
