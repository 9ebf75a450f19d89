set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^init\|^/root/ref" | tail -25 > gpurun_out/r01b_tests.log
cat gpurun_out/r01b_tests.log | tail -12
python tools/profile_solve.py --max-iter 2 --batch 65536 2>&1 | tail -1
python tools/profile_solve.py --max-iter 6 --batch 65536 2>&1 | tail -1
python tools/profile_solve.py --max-iter 0 --batch 65536 2>&1 | tail -1
python tools/profile_solve.py --max-iter 0 --batch 37888 2>&1 | tail -1
python bench.py --steps 2 --warmup 1 --no-cpu 2>&1 | tail -1 > gpurun_out/r01b_bench_quad_first.json
cat gpurun_out/r01b_bench_quad_first.json
python bench.py --workload mpc_dcmotor --steps 2 --warmup 1 --no-cpu 2>&1 | tail -1 > gpurun_out/r01b_bench_dcmotor_first.json
cat gpurun_out/r01b_bench_dcmotor_first.json
timeout 600 ncu --set full --clock-control none -k regex:tcm_solve -c 1 -o gpurun_out/r01b_quad_compiled python tools/profile_solve.py --batch 37888 --max-iter 2 2>&1 | tail -3
