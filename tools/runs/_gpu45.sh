cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --parity 4096 --no-strong "$@" > gpurun_out/r03u_$n.json 2> gpurun_out/r03u_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03u_$n.json"))
    p=d["parity"]
    print("$n value %.0f us/step %.4f steps/solve %.3f hist %s | parity ok=%s status_mismatches=%s iter_equal=%s within1=%s beyond1=%s maxdiff=%s max_rel_err=%.3g"%(d["value"],d["us_per_newton_step"],d["newton_steps_per_solve"],d["status_histogram"],p["ok"],p["status_mismatches"],p["iter_equal"],p["iter_within_1"],p["iter_beyond_1"],p["max_abs_iter_diff"],p["max_rel_output_err"]))
except Exception as e:
    print("$n FAILED", e)
PY
  grep -v "attached" gpurun_out/r03u_$n.err | tail -2
}
run quad_fmad --mopts fmad=1
run l1l2_fmad --workload l1l2 --mopts fmad=1
run dcmotor_fmad --workload mpc_dcmotor --mopts fmad=1
run sls_fmad --workload sls --mopts fmad=1
