cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r03z_bench_quadcopter.json 2> gpurun_out/r03z_bench_quadcopter.err
python - <<PY
import json
d=json.load(open("gpurun_out/r03z_bench_quadcopter.json"))
print("quad value",d["value"],"ms/step",d["ms_per_step"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"],"cpu",d["cpu_baseline"]["value"],d["cpu_baseline"]["cores"],"parity",(d["parity"] or {}).get("ok"),(d["parity"] or {}).get("checked"), d["config"]["pipeline"])
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03z_bench_reference.json 2> /dev/null
head -c 300 gpurun_out/r03z_bench_reference.json; echo
timeout 900 python bench.py --workload dense_ls --steps 5 --warmup 3 > gpurun_out/r03z_bench_dense_ls.json 2> gpurun_out/r03z_bench_dense_ls.err
python - <<PY
import json
d=json.load(open("gpurun_out/r03z_bench_dense_ls.json"))
print("dense value",d["value"],"ms/step",d["ms_per_step"],"roof",d["roofline"]["kernel"],d["roofline"]["achieved"],d["roofline"]["frac"],"e2e",d["e2e"]["value"],"cpu",d["cpu_baseline"]["value"],"parity",d["parity"]["ok"], {k:(v["ms"],v["tflops"]) for k,v in d["roofline"]["kernels"].items()})
PY
run() { # name, args
  n=$1; shift
  timeout 900 python bench.py --steps 20 --warmup 4 --no-strong "$@" > gpurun_out/r03z_bench_$n.json 2> gpurun_out/r03z_bench_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r03z_bench_$n.json"))
    print("$n value %.0f ms/step %.1f frac %.3f e2e %.0f cpu %.0f parity %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["cpu_baseline"]["value"],(d["parity"] or {}).get("ok")))
except Exception as e:
    print("$n FAILED", e)
PY
}
run l1l2 --workload l1l2
run mpc_dcmotor --workload mpc_dcmotor
run sls --workload sls
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r03z_launches_bench_quadcopter.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --parity 0 --no-strong > gpurun_out/ncu_b.log 2>&1
tail -1 gpurun_out/ncu_b.log | cut -c1-200
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r03w_launches_p1.csv python bench.py --steps 1 --warmup 1 --pipeline 1 --no-cpu --no-e2e --parity 0 --no-strong > gpurun_out/ncu_c.log 2>&1
tail -1 gpurun_out/ncu_c.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tcm_solve_kernel -s 3 -c 1 -o /tmp/r03z_quad_full python tools/profile_solve.py --batch 56832 --max-iter 8 > gpurun_out/ncu_d.log 2>&1
ncu -i /tmp/r03z_quad_full.ncu-rep --page raw --csv > /tmp/raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/raw.csv > gpurun_out/r03z_ncu_full_tcm_solve_kernel_summary.txt
ncu -i /tmp/r03z_quad_full.ncu-rep --page details > gpurun_out/r03z_ncu_full_tcm_solve_kernel_details.txt 2>/dev/null
ncu -i /tmp/r03z_quad_full.ncu-rep --page source --csv > /tmp/src.csv 2>/dev/null
python tools/ncu_source_top.py /tmp/src.csv > gpurun_out/r03z_ncu_source_top.txt 2>&1
ls -la gpurun_out/r03z* gpurun_out/r03w*
