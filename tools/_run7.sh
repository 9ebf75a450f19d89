python bench.py --steps 6 --warmup 3 2>&1 | tail -1 > gpurun_out/r01g_bench_quad_pipeline3.json; cat gpurun_out/r01g_bench_quad_pipeline3.json
python bench.py --steps 3 --warmup 3 --pipeline 1 --no-cpu 2>&1 | tail -1 > gpurun_out/r01g_bench_quad_pipeline1.json; cat gpurun_out/r01g_bench_quad_pipeline1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01g_launches_bench_quad.csv python bench.py --steps 2 --warmup 1 --no-cpu --pipeline 1 > /dev/null 2>&1
tail -5 gpurun_out/r01g_launches_bench_quad.csv
timeout 600 ncu --set full --clock-control none -k regex:tcm_solve -c 1 -o /tmp/repfull python tools/profile_solve.py --batch 37888 --max-iter 2 2>&1 | tail -1
ncu -i /tmp/repfull.ncu-rep --page raw --csv > gpurun_out/r01g_ncu_full_quad.csv
ncu -i /tmp/repfull.ncu-rep --page details > gpurun_out/r01g_ncu_full_quad_details.txt
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 > gpurun_out/r01g_bench_reference.json; cat gpurun_out/r01g_bench_reference.json
