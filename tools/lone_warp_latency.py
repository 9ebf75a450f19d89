"""Measurement tool: Newton-iteration latency of a lone warp (the straggler phases of a batch), warm caches, nothing
else on the GPU -- what the drain at the end of a bench run costs per iteration.  python tools/lone_warp_latency.py"""
import json
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from tenscalc_b200 import problems

out = {}
gen, args = problems.mpc_quadcopter()
for backend, batch in (("compiled", 20),) if "--loaded" in sys.argv else (("compiled", 20), ("compiled", 32), ("compiled", 128), ("tape", 20)):
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

# the same 20 instances while two full batches (65 536 instances each, own handles and streams) run their bulk
# phases next to them: the condition of the straggler phases in the steady state of bench.py
if "--loaded" in sys.argv:
    big = []
    for k in range(2):
        B = 65536
        Pb, Vb = problems.synthetic_inputs("mpc_quadcopter", gen, B, seed=1 + k)
        o = gen(B)
        for n, _ in gen.parameters:
            getattr(o, "setP_" + n)(Pb[n])
        for n, _ in gen.variables:
            getattr(o, "setV_" + n)(Vb[n])
        big.append((o, np.zeros(B, np.int32), np.zeros(B, np.int32)))
    P, V = problems.synthetic_inputs("mpc_quadcopter", gen, 20, seed=0)
    small = gen(20)
    for n, _ in gen.parameters:
        getattr(small, "setP_" + n)(P[n])
    for n, _ in gen.variables:
        getattr(small, "setV_" + n)(V[n])
    st, it, t0 = small.solve(args["mu0"], args["maxIter"], -1)            # warm, alone
    for n, _ in gen.variables:
        getattr(small, "setV_" + n)(V[n])
    for o, s_, i_ in big:
        o.solve_async(s_, i_, args["mu0"], args["maxIter"])
    st, it, t1 = small.solve(args["mu0"], args["maxIter"], -1)            # next to the two bulk solves
    for o, s_, i_ in big:
        o.synchronize()
    out["compiled_b20_beside_two_bulk_batches"] = dict(
        max_iter=int(it.max()), alone_ms=1e3 * t0, loaded_ms=1e3 * t1, ms_per_newton_iteration=1e3 * t1 / int(it.max()))
    print(out["compiled_b20_beside_two_bulk_batches"], flush=True)
json.dump(out, open("gpurun_out/" + ("r04i_lone_warp_loaded.json" if "--loaded" in sys.argv else "r04h_lone_warp_latency.json"), "w"), indent=1)
