cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r04g_gpu_tests.log 2>&1; tail -3 gpurun_out/r04g_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r04g_bench_quadcopter.json 2> gpurun_out/r04g_bench_quadcopter.err
python - <<PY
import json
d=json.load(open("gpurun_out/r04g_bench_quadcopter.json"))
print("quad value",d["value"],"ms/step",d["ms_per_step"],"frac",d["roofline"]["frac"],"traffic",d["roofline"]["traffic"],"e2e",d["e2e"]["value"],"cpu",d["cpu_baseline"]["value"],d["cpu_baseline"]["cores"],"parity",(d["parity"] or {}).get("ok"),(d["parity"] or {}).get("checked"))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r04g_bench_reference.json 2> /dev/null
head -c 200 gpurun_out/r04g_bench_reference.json; echo
