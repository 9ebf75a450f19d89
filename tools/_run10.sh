nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 2>&1 | tail -1 > gpurun_out/r01i_bench_quadcopter_2gpu.json; cat gpurun_out/r01i_bench_quadcopter_2gpu.json | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-300
