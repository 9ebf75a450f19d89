cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "tenscalc_b200: module" | tail -6 > gpurun_out/r02y_gpu_tests.log
cat gpurun_out/r02y_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu > gpurun_out/r02y_bench_quadcopter_2gpu.json 2> gpurun_out/r02y_2gpu.err
tail -2 gpurun_out/r02y_2gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r02y_bench_quadcopter_2gpu.json"))
print("2gpu value",d["value"],"ms/step",d["ms_per_step"],"e2e",d["e2e"],"strong",d.get("strong_scaling"),"per_rank",d["per_rank"]["total_ms"])
PY
