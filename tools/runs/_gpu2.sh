set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for d in 3 5 8; do
timeout 600 python bench.py --steps 20 --warmup 5 --pipeline $d --no-cpu --no-e2e --parity 0 > gpurun_out/r02b_bench_quad_p$d.json 2> gpurun_out/r02b_err_p$d.txt
cut -c1-330 gpurun_out/r02b_bench_quad_p$d.json; tail -2 gpurun_out/r02b_err_p$d.txt
done
python tools/profile_solve.py --batch 37888 --max-iter 6 --repeat 2 2>&1 | tail -2
python tools/profile_solve.py --batch 75776 --max-iter 6 --repeat 2 2>&1 | tail -2
python tools/profile_solve.py --batch 65536 --max-iter 0 --repeat 2 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tcm_solve_kernel -s 1 -c 1 -o gpurun_out/r02b_quad_full python tools/profile_solve.py --batch 37888 --max-iter 2 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out
