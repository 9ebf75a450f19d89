python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^init\|^/root/ref" | tail -12 | tee gpurun_out/r01h_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 6 --warmup 3 2>&1 | tail -1 > gpurun_out/r01h_bench_quadcopter.json; cat gpurun_out/r01h_bench_quadcopter.json | cut -c1-400
python bench.py --steps 8 --warmup 3 --pipeline 4 --no-cpu 2>&1 | tail -1 > gpurun_out/r01h_bench_quadcopter_p4.json; cat gpurun_out/r01h_bench_quadcopter_p4.json | cut -c1-200
for w in mpc_dcmotor l1l2 sls; do python bench.py --workload $w --steps 6 --warmup 3 2>&1 | tail -1 > gpurun_out/r01h_bench_$w.json; cat gpurun_out/r01h_bench_$w.json | cut -c1-250; done
