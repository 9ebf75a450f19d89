cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r03z_bench_quadcopter.json 2> gpurun_out/r03z_bench_quadcopter.err
python - <<PY
import json
d=json.load(open("gpurun_out/r03z_bench_quadcopter.json"))
print("quad value",d["value"],"ms/step",d["ms_per_step"],"frac",d["roofline"]["frac"],"traffic",d["roofline"]["traffic"],"e2e",d["e2e"]["value"],"cpu",d["cpu_baseline"]["value"],d["cpu_baseline"]["cores"],"parity",(d["parity"] or {}).get("ok"),(d["parity"] or {}).get("checked"))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03z_bench_reference.json 2> /dev/null
head -c 200 gpurun_out/r03z_bench_reference.json; echo
for w in l1l2 mpc_dcmotor sls; do
timeout 900 python bench.py --steps 20 --warmup 4 --no-strong --workload $w > gpurun_out/r03z_bench_$w.json 2> gpurun_out/r03z_bench_$w.err
python - <<PY
import json
d=json.load(open("gpurun_out/r03z_bench_$w.json"))
print("$w value %.0f ms/step %.1f frac %.3f e2e %.0f cpu %.0f parity %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["cpu_baseline"]["value"],(d["parity"] or {}).get("ok")))
PY
done
timeout 900 python bench.py --workload dense_ls --steps 5 --warmup 3 > gpurun_out/r03z_bench_dense_ls.json 2> gpurun_out/r03z_bench_dense_ls.err
python - <<PY
import json
d=json.load(open("gpurun_out/r03z_bench_dense_ls.json"))
print("dense value",d["value"],"ms/step",d["ms_per_step"],"e2e",d["e2e"]["value"],"parity",d["parity"]["ok"])
PY
