cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 14 --warmup 4 --no-cpu --no-e2e --parity 256 --no-strong "$@" > gpurun_out/r03r_$n.json 2> gpurun_out/r03r_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03r_$n.json"))
    print("$n value %.0f ms/step %.1f us/step %.4f frac %.3f parity %s"%(d["value"],d["ms_per_step"],d["us_per_newton_step"],d["roofline"]["frac"],(d["parity"] or {}).get("ok")), d["config"]["pipeline"][:10])
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03r_$n.err | tail -3
}
run b56832p3 --batch 56832 --pipeline 3
run b56832p7 --batch 56832 --pipeline 7
run b60000p7 --batch 60000 --pipeline 7
run b65536p7 --pipeline 7
run b49152p7 --batch 49152 --pipeline 7
