cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(cd tools/bench && ./fp64peak) | tee gpurun_out/fp64_peak.json
timeout 900 python -m pytest tests/test_dense.py -m gpu -x -q -s 2>&1 | tail -8 > gpurun_out/r02l_dense_tests.log
cat gpurun_out/r02l_dense_tests.log
timeout 600 python tools/dense_probe.py --batch 1024 2>&1 | tail -4 | tee gpurun_out/r02l_dense_probe.txt
timeout 900 python bench.py --workload dense_ls --steps 5 --warmup 3 > gpurun_out/r02l_bench_dense_ls.json 2> gpurun_out/r02l_bench_dense_ls.err
tail -3 gpurun_out/r02l_bench_dense_ls.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02l_bench_dense_ls.json"))
print("dense value",d["value"],"ms/step",d["ms_per_step"],"roof",d["roofline"]["achieved"],d["roofline"]["frac"],d["roofline"]["kernels"],"e2e",d["e2e"],"cpu",d["cpu_baseline"],"parity",d["parity"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tcd_ -c 2 -o gpurun_out/r02l_dense python tools/dense_probe.py --batch 296 --repeat 1 > gpurun_out/ncu_dense.log 2>&1
ncu -i gpurun_out/r02l_dense.ncu-rep --page raw --csv > gpurun_out/r02l_dense_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_dense.log
