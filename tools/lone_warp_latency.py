"""Measurement tool: Newton-iteration latency of a lone warp (the straggler phases of a batch), warm caches, nothing
else on the GPU -- what the drain at the end of a bench run costs per iteration.  python tools/lone_warp_latency.py"""
import json
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from tenscalc_b200 import problems

out = {}
gen, args = problems.mpc_quadcopter()
for backend, batch in (("compiled", 20), ("compiled", 32), ("compiled", 128), ("tape", 20)):
    P, V = problems.synthetic_inputs("mpc_quadcopter", gen, batch, seed=0)
    obj = gen(batch, backend=backend)
    times = []
    for rep in range(3):
        for n, _ in gen.parameters:
            getattr(obj, "setP_" + n)(P[n])
        for n, _ in gen.variables:
            getattr(obj, "setV_" + n)(V[n])
        st, it, t = obj.solve(args["mu0"], args["maxIter"], -1)
        times.append(t)
    obj.close()
    key = f"{backend}_b{batch}"
    out[key] = dict(batch=batch, max_iter=int(it.max()), mean_iter=float(it.mean()), solve_ms=[1e3 * x for x in times],
                    ms_per_newton_iteration=1e3 * min(times) / int(it.max()))
    print(key, out[key], flush=True)
json.dump(out, open("gpurun_out/r04h_lone_warp_latency.json", "w"), indent=1)
