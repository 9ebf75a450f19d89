cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r04f_bench_quadcopter_2gpu.json 2> gpurun_out/r04f_2gpu.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r04f_bench_quadcopter_2gpu.json") if l.startswith("{")][-1])
print("2gpu value",d["value"],"ms/step",d["ms_per_step"],"e2e",d["e2e"]["value"],"scaling",d["scaling"],"strong",(d.get("strong_scaling") or {}).get("value"),"parity",(d["parity"] or {}).get("ok"))
PY
tail -3 gpurun_out/r04f_2gpu.err
