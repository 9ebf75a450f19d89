cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(layout_options and roll) or componentwise or isa_probe" > gpurun_out/r03v_roll_tests.log 2>&1; tail -3 gpurun_out/r03v_roll_tests.log
run() { # name, args
  n=$1; shift
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --parity 4096 --no-strong "$@" > gpurun_out/r03v_$n.json 2> gpurun_out/r03v_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03v_$n.json"))
    p=d["parity"]
    print("$n value %.0f us/step %.4f steps/solve %.3f | parity ok=%s status_mismatches=%s iter_equal=%s max_rel_err=%.3g"%(d["value"],d["us_per_newton_step"],d["newton_steps_per_solve"],p["ok"],p["status_mismatches"],p["iter_equal"],p["max_rel_output_err"]))
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03v_$n.err | tail -2
}
run quad_roll50 --mopts roll=50
run l1l2_roll50 --workload l1l2 --mopts roll=50
run quad_default
