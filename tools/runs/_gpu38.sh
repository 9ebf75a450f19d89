cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r03n_bench_quadcopter.json 2> gpurun_out/r03n.err
python - <<PY
import json
d=json.load(open("gpurun_out/r03n_bench_quadcopter.json"))
print("quad value %.0f ms/step %.1f frac %.3f e2e %.0f e2e ms/step %.1f parity %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["parity"]))
print(d["e2e"]["api"])
PY
grep -v attached gpurun_out/r03n.err | tail -3; free -g | head -2
