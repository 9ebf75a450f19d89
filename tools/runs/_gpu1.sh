set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
tools/bench/fp64peak > gpurun_out/fp64_peak.json 2>&1
cat gpurun_out/fp64_peak.json
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_gpu_tests.log
cat gpurun_out/r02a_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench_quadcopter.json 2> gpurun_out/r02a_bench_quadcopter.err
cat gpurun_out/r02a_bench_quadcopter.json
tail -5 gpurun_out/r02a_bench_quadcopter.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tcm_solve_kernel -c 3 --csv --log-file gpurun_out/r02a_ncu_dram_quad.csv python bench.py --steps 1 --warmup 1 --pipeline 1 --no-cpu --no-e2e --parity 0 > gpurun_out/ncu_bench.log 2>&1
tail -8 gpurun_out/r02a_ncu_dram_quad.csv
