cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02w_launches_p1.csv python bench.py --steps 1 --warmup 1 --pipeline 1 --no-cpu --no-e2e --parity 0 --no-strong > gpurun_out/ncu_b.log 2>&1
tail -1 gpurun_out/ncu_b.log | cut -c1-300
for mi in 2 6 12 20 28; do python tools/profile_solve.py --batch 56832 --max-iter $mi 2>&1 | tail -1 | cut -c1-200; done
