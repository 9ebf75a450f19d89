cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "tenscalc_b200: module" | tail -6 > gpurun_out/r02u_gpu_tests.log
cat gpurun_out/r02u_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02u_bench_quadcopter.json 2> gpurun_out/r02u_bench_quadcopter.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02u_bench_quadcopter.json"))
print("quad value",d["value"],"ms/step",d["ms_per_step"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"],"cpu",d["cpu_baseline"]["value"],d["cpu_baseline"]["cores"],"parity",(d["parity"] or {}).get("ok"),(d["parity"] or {}).get("checked"))
PY
timeout 900 python bench.py --workload dense_ls --steps 5 --warmup 3 > gpurun_out/r02u_bench_dense_ls.json 2> gpurun_out/r02u_bench_dense_ls.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02u_bench_dense_ls.json"))
print("dense value",d["value"],"ms/step",d["ms_per_step"],"roof",d["roofline"]["kernel"],d["roofline"]["achieved"],d["roofline"]["frac"],"e2e",d["e2e"]["value"],"cpu",d["cpu_baseline"]["value"],"parity",d["parity"]["ok"])
PY
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02u_launches_bench_quadcopter.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --parity 0 --no-strong > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log | cut -c1-200
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tcm_solve --csv --log-file gpurun_out/r02u_ncu_dram_quad.csv python bench.py --steps 1 --warmup 0 --pipeline 1 --no-cpu --no-e2e --parity 0 --no-strong > gpurun_out/ncu_c.log 2>&1
tail -1 gpurun_out/ncu_c.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tcm_solve_kernel -s 1 -c 1 -o gpurun_out/r02u_quad_full python tools/profile_solve.py --batch 56832 --max-iter 2 > gpurun_out/ncu_d.log 2>&1
ncu -i gpurun_out/r02u_quad_full.ncu-rep --page raw --csv > /tmp/raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/raw.csv > gpurun_out/r02u_ncu_full_tcm_solve_kernel_summary.txt
grep -v pcsamp gpurun_out/r02u_ncu_full_tcm_solve_kernel_summary.txt | head -50
