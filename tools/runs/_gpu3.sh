set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for mi in 6 12 20 28 36 50; do python tools/profile_solve.py --batch 37888 --max-iter $mi 2>&1 | tail -1 | cut -c1-120; done
python tools/profile_solve.py --batch 56832 --max-iter 6 --warps-per-sm 12 --max-loads 64 --repeat 2 2>&1 | tail -1 | cut -c1-150
python tools/profile_solve.py --batch 75776 --max-iter 6 --warps-per-sm 16 --max-loads 44 --repeat 2 2>&1 | tail -1 | cut -c1-150
python tools/profile_solve.py --batch 65536 --max-iter 0 --warps-per-sm 12 --max-loads 64 --repeat 2 2>&1 | tail -1 | cut -c1-150
python tools/profile_solve.py --batch 65536 --max-iter 0 --warps-per-sm 16 --max-loads 44 --repeat 2 2>&1 | tail -1 | cut -c1-150
python tools/lane_occupancy.py > gpurun_out/r02c_lane_occupancy_warp.txt 2>&1; tail -30 gpurun_out/r02c_lane_occupancy_warp.txt
timeout 900 ncu --set full --clock-control none -k regex:tcm_solve_kernel -s 1 -c 1 -o /tmp/r02c_quad_full python tools/profile_solve.py --batch 37888 --max-iter 2 > gpurun_out/ncu3.log 2>&1
ncu -i /tmp/r02c_quad_full.ncu-rep --page raw --csv > /tmp/raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/raw.csv > gpurun_out/r02c_ncu_full_tcm_solve_kernel_summary.txt
ncu -i /tmp/r02c_quad_full.ncu-rep --page details > gpurun_out/r02c_ncu_full_tcm_solve_kernel_details.txt 2>/dev/null
cat gpurun_out/r02c_ncu_full_tcm_solve_kernel_summary.txt
