cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu --no-e2e --no-strong --parity 0 "$@" > gpurun_out/r02h_$n.json 2> gpurun_out/r02h_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02h_$n.json"))
    print("$n value %.0f ms/step %.1f frac %.3f launches %d"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["gpu_launches"]))
except Exception as e:
    print("$n FAILED", e)
PY
}
run base_p4 --pipeline 4
run ls4w16 --pipeline 4 --mopts lockstep=4,warps_per_sm=16,max_loads=44
run ls8w24 --pipeline 4 --mopts lockstep=8,warps_per_sm=24,max_loads=28
run ls2w12 --pipeline 4 --mopts lockstep=2,warps_per_sm=12,max_loads=64
run ls1w12 --pipeline 4 --mopts lockstep=1,warps_per_sm=12,max_loads=64
run ls4w12ml72 --pipeline 4 --mopts lockstep=4,warps_per_sm=12,max_loads=72
run ls4w12_p5 --pipeline 5 --mopts lockstep=4,warps_per_sm=12,max_loads=64
run ls4w12_p3 --pipeline 3 --mopts lockstep=4,warps_per_sm=12,max_loads=64
