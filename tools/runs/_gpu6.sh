cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "tenscalc_b200: module" | tail -12 > gpurun_out/r02e_gpu_tests.log
cat gpurun_out/r02e_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --parity 4096 > gpurun_out/r02e_bench_quad_phased.json 2> gpurun_out/r02e_err.txt
python - <<PY
import json
d=json.load(open("gpurun_out/r02e_bench_quad_phased.json"))
print("phased value",d["value"],"ms/step",d["ms_per_step"],"parity",d["parity"],"frac",d["roofline"]["frac"],"launches",d["gpu_launches"])
PY
python tools/lane_occupancy.py 2>&1 | grep -v module | head -8
