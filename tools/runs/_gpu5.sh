cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "tenscalc_b200: module" | tail -25 > gpurun_out/r02d_gpu_tests.log
cat gpurun_out/r02d_gpu_tests.log
for ph in 1 0; do
TC_PHASED=$ph timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --parity 2048 > gpurun_out/r02d_bench_quad_phased$ph.json 2> gpurun_out/r02d_err$ph.txt
cut -c1-420 gpurun_out/r02d_bench_quad_phased$ph.json; grep -v "module" gpurun_out/r02d_err$ph.txt | tail -3
python - <<PY
import json
d=json.load(open("gpurun_out/r02d_bench_quad_phased$ph.json"))
print("phased=$ph value",d["value"],"ms/step",d["ms_per_step"],"parity",d["parity"]["ok"],d["parity"]["iter_equal"],"frac",d["roofline"]["frac"],"launches",d["gpu_launches"])
PY
done
TC_PHASED=1 python tools/profile_solve.py --batch 65536 --max-iter 0 --repeat 2 2>&1 | tail -1 | cut -c1-130
