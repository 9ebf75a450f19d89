cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "tenscalc_b200: module" | tail -15 > gpurun_out/r02j_gpu_tests.log
cat gpurun_out/r02j_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench_quadcopter.json 2> gpurun_out/r02j_bench_quadcopter.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02j_bench_quadcopter.json"))
print("quad value",d["value"],"ms/step",d["ms_per_step"],"frac",d["roofline"]["frac"],"e2e",d["e2e"],"cpu",d["cpu_baseline"]["value"],"parity",(d["parity"] or {}).get("ok"),(d["parity"] or {}).get("checked"))
PY
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu --no-e2e --no-strong --parity 0 "$@" > gpurun_out/r02j_$n.json 2> gpurun_out/r02j_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02j_$n.json"))
    print("$n value %.0f ms/step %.1f frac %.3f launches %d"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["gpu_launches"]))
except Exception as e:
    print("$n FAILED", e)
PY
}
run l1l2_new --workload l1l2
run l1l2_old --workload l1l2 --mopts lockstep=1,warps_per_sm=8,max_loads=96
run dcmotor_base --workload mpc_dcmotor
run dcmotor_ls4 --workload mpc_dcmotor --mopts lockstep=4
run dcmotor_ls8 --workload mpc_dcmotor --mopts lockstep=8
run sls_base --workload sls
run sls_ls4 --workload sls --mopts lockstep=4
run sls_ls8 --workload sls --mopts lockstep=8
