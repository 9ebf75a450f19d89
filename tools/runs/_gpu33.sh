cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --parity 512 --no-strong "$@" > gpurun_out/r03i_$n.json 2> gpurun_out/r03i_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03i_$n.json"))
    print("$n value %.0f ms/step %.1f us/step %.4f frac %.3f e2e %.0f parity %s"%(d["value"],d["ms_per_step"],d["us_per_newton_step"],d["roofline"]["frac"],d["e2e"]["value"],(d["parity"] or {}).get("ok")))
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03i_$n.err | tail -3
  nvidia-smi --query-gpu=memory.used --format=csv,noheader
}
run p5 --mopts prefetch=0
run p7 --mopts prefetch=0 --pipeline 7
run p8 --mopts prefetch=0 --pipeline 8
