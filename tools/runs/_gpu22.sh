cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
free -g | head -2; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu --no-strong > gpurun_out/r02y_bench_quadcopter_2gpu_weak.json 2> gpurun_out/r02y_2gpu_weak.err
echo "exit code weak: $?"
grep -v "attached (phased" gpurun_out/r02y_2gpu_weak.err | tail -15 | cut -c1-300
head -c 600 gpurun_out/r02y_bench_quadcopter_2gpu_weak.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu > gpurun_out/r02y_bench_quadcopter_2gpu.json 2> gpurun_out/r02y_2gpu.err
echo "exit code full: $?"
grep -v "attached (phased" gpurun_out/r02y_2gpu.err | tail -15 | cut -c1-300
dmesg 2>/dev/null | tail -5
head -c 600 gpurun_out/r02y_bench_quadcopter_2gpu.json; echo
