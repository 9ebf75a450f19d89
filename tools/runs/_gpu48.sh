cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for w in l1l2 mpc_dcmotor sls; do
timeout 600 python bench.py --steps 20 --warmup 4 --no-strong --workload $w > gpurun_out/r04c_bench_$w.json 2> gpurun_out/r04c_bench_$w.err
python - <<PY
import json
d=json.load(open("gpurun_out/r04c_bench_$w.json"))
print("$w value %.0f ms/step %.1f frac %.3f e2e %.0f cpu %.0f parity %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["cpu_baseline"]["value"],(d["parity"] or {}).get("ok")))
PY
done
timeout 600 python bench.py --workload dense_ls --steps 5 --warmup 3 > gpurun_out/r04c_bench_dense_ls.json 2> gpurun_out/r04c_bench_dense_ls.err
python - <<PY
import json
d=json.load(open("gpurun_out/r04c_bench_dense_ls.json"))
print("dense value",d["value"],"ms/step",d["ms_per_step"],"e2e",d["e2e"]["value"],"parity",d["parity"]["ok"])
PY
