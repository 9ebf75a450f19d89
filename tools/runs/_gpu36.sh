cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 14 --warmup 4 --no-cpu --no-e2e --parity 256 --no-strong "$@" > gpurun_out/r03m_$n.json 2> gpurun_out/r03m_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03m_$n.json"))
    print("$n value %.0f ms/step %.1f us/step %.4f frac %.3f parity %s launches %s"%(d["value"],d["ms_per_step"],d["us_per_newton_step"],d["roofline"]["frac"],(d["parity"] or {}).get("ok"),d["gpu_launches"]))
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03m_$n.err | tail -3
}
run default
run ph1x40 --phases 1x40,2x12,4x16,8x16,16x16,32
run ph1x24 --phases 1x24,2x20,4x16,8x16,16x16,32
run ph1_3x --phases 1,3x21,4x16,8x16,16x16,32
