python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^init\|^/root/ref" | tail -12 | tee gpurun_out/r01k_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
P="python tools/profile_solve.py --max-iter 0 --repeat 2"
$P --batch 1048576 --workload sls 2>&1 | tail -1 | cut -c1-110
$P --batch 1048576 --workload sls --warps-per-sm 8 --max-loads 96 2>&1 | tail -1 | cut -c1-110
python bench.py --steps 6 --warmup 3 2>&1 | tail -1 > gpurun_out/r01k_bench_quadcopter.json; cat gpurun_out/r01k_bench_quadcopter.json | cut -c1-300
for w in mpc_dcmotor l1l2 sls; do python bench.py --workload $w --steps 6 --warmup 3 2>&1 | tail -1 > gpurun_out/r01k_bench_$w.json; cat gpurun_out/r01k_bench_$w.json | cut -c1-200; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01k_launches_bench_quadcopter.csv python bench.py --steps 2 --warmup 1 --no-cpu --pipeline 1 > /dev/null 2>&1
tail -3 gpurun_out/r01k_launches_bench_quadcopter.csv | cut -c1-200
