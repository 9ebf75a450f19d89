cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu --no-e2e --no-strong --parity 0 "$@" > gpurun_out/r02i_$n.json 2> gpurun_out/r02i_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02i_$n.json"))
    print("$n value %.0f ms/step %.1f frac %.3f launches %d"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["gpu_launches"]))
except Exception as e:
    print("$n FAILED", e)
PY
}
run ls4w12ml72_p5 --pipeline 5 --mopts lockstep=4,warps_per_sm=12,max_loads=72
run ls4w12ml72_p6 --pipeline 6 --mopts lockstep=4,warps_per_sm=12,max_loads=72
run ls4w12ml72_p8 --pipeline 8 --mopts lockstep=4,warps_per_sm=12,max_loads=72
run ls4w12ml80_p5 --pipeline 5 --mopts lockstep=4,warps_per_sm=12,max_loads=80
run ls2w12ml72_p5 --pipeline 5 --mopts lockstep=2,warps_per_sm=12,max_loads=72
run ls4w16ml52_p5 --pipeline 5 --mopts lockstep=4,warps_per_sm=16,max_loads=52
run ls4w16ml44_p5 --pipeline 5 --mopts lockstep=4,warps_per_sm=16,max_loads=44
