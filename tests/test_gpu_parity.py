"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle
(reference driver + reference emitter, prebuilt in oracle/_ref) on the same seeded inputs.

Bar (BASELINE.json north_star): identical exit status, iteration count within +-1,
outputs within 1e-8 relative.  For problems without libm calls the kernels reproduce the
oracle's rounding sequence, so iterations are compared exactly and outputs to 1e-12.
"""
import numpy as np
import pytest

from conftest import generated
from tenscalc_b200 import problems

pytestmark = pytest.mark.gpu

# extra emitter options for the compiled back end, e.g. TC_MODULE_OPTS='{"driver": "lanes"}'
MODULE_OPTS = __import__("json").loads(__import__("os").environ.get("TC_MODULE_OPTS", "{}"))

AFFINE = {"algorithm": "ipmPD_CS", "skipAffine": False}     # Mehrotra predictor (ipmPD_CSsolver.c:576-665; builder ipmPD_CS.m)
CASES = [("sls", 512, {}), ("mpc_dcmotor", 512, {}), ("l1l2", 128, {}), ("mpc_quadcopter", 96, {}),
         ("mpc_quadcopter", 256, {"T_": 20, "useInertia": True}),
         ("sls", 512, AFFINE), ("mpc_dcmotor", 512, AFFINE), ("mpc_dcmotor", 512, {**AFFINE, "smallerNewtonMatrix": True}),
         ("mpc_quadcopter", 256, {"T_": 20, **AFFINE}),
         # every unary opcode of the ISA (instructionsTableUTHash.c:1347-1539) and multi-line sums of products
         ("isa_probe", 512, {}),
         # Nash equilibria: cmex2equilibriumLatentCS -> ipmPDeq_CSsolver (lib/ipmPDeq_CSsolver.c, examples/TCgames.m),
         # non-symmetric Newton matrix factored by the fixed-pattern LU; oracle = the reference's unmodified driver
         ("matrix_game", 512, {}), ("qsaddle", 512, {}), ("qsaddle_latent", 512, {}),
         ("qsaddle", 256, {"useLDL": True}),                                  # normal equations + LDL' (ipmPDeqlat_CS.m:645-651)
         ("matrix_game", 256, {"skipAffine": False}),
         ("qsaddle", 256, {"addEye2Hessian": True, "adjustAddEye2Hessian": True}),     # regulariser retries (:318-374)
         # reduced Newton matrix of the equilibrium builder (ipmPDeqlat_CS.m:314-561), also with latent variables and Mehrotra
         ("qsaddle", 256, {"smallerNewtonMatrix": True}), ("qsaddle_latent", 256, {"smallerNewtonMatrix": True}),
         ("matrix_game", 256, {"smallerNewtonMatrix": True, "skipAffine": False}),
         # min-max: cmex2minmaxCS -> ipmPDminmax_CSsolver (three regularisers, inertia control).  PARITY UNPINNED:
         # the oracle's driver is a C transcription of the reference's MATLAB solver (oracle/ipm_driver_minmax_port.c)
         ("minmax_box", 512, {}), ("minmax_box_latent", 512, {}), ("minmax_box", 256, {"N": 16})]                         # affine-scaling step of the eq driver (:326-398)


def _oracle(gen, name):
    from oracle.oracle import Oracle
    return Oracle(gen, "ref")


@pytest.mark.parametrize("backend", ["compiled", "tape"])
@pytest.mark.parametrize("name,batch,kw", CASES)
def test_parity_with_oracle(name, batch, kw, backend):
    gen, args = generated(name, **kw)
    P, V = problems.synthetic_inputs(name, gen, batch, seed=1)
    st_o, it_o, out_o, _ = _oracle(gen, name).solve_batch(P, V, args["mu0"], args["maxIter"])
    obj = gen(batch, backend=backend, **(MODULE_OPTS if backend == "compiled" else {}))
    assert obj.has_module() == (backend == "compiled")
    for n, _ in gen.parameters:
        getattr(obj, "setP_" + n)(P[n])
    for n, _ in gen.variables:
        getattr(obj, "setV_" + n)(V[n])
    st, it, t = obj.solve(args["mu0"], args["maxIter"], -1)
    outs = obj.getOutputs()
    if not isinstance(outs, dict):
        outs = {n: o for (n, _), o in zip(gen.outputs, outs)}
    obj.close()
    assert np.array_equal(st, st_o), f"status mismatch: {np.nonzero(st != st_o)[0][:10]}"
    assert np.abs(it.astype(int) - it_o.astype(int)).max() <= 1
    ok = st_o == 0
    assert ok.mean() > 0.5
    for k in out_o:
        a, b = outs[k][..., ok], out_o[k][..., ok]
        scale = np.maximum(np.abs(b).max(axis=tuple(range(b.ndim - 1)), keepdims=True), 1e-300)
        assert np.max(np.abs(a - b) / scale) <= 1e-8, k
    print(f"{name}[{backend}]: B={batch} iters {it.min()}..{it.max()} equal_iters={np.array_equal(it, it_o)} "
          f"batch time {t * 1e3:.2f} ms")


def _solve(gen, args, P, V, batch, **opts):
    opts = {k: v for k, v in opts.items() if not (k == "driver" and v == "warp")}   # the default driver
    obj = gen(batch, **opts)
    for n, _ in gen.parameters:
        getattr(obj, "setP_" + n)(P[n])
    for n, _ in gen.variables:
        getattr(obj, "setV_" + n)(V[n])
    return obj


def test_async_solve_and_two_handles_in_flight():
    """tc_solve_async on two handles (own stream and scratchbook each) gives the results of tc_solve"""
    gen, args = generated("sls")
    batch = 2048
    P, V = problems.synthetic_inputs("sls", gen, batch, seed=5)
    ref = _solve(gen, args, P, V, batch)
    st0, it0, _ = ref.solve(args["mu0"], args["maxIter"], -1)
    y0 = ref.getOutputs()
    ref.close()
    objs = [_solve(gen, args, P, V, batch) for _ in range(2)]
    res = [(np.full(batch, -1, np.int32), np.full(batch, -1, np.int32)) for _ in objs]
    for o, (st, it) in zip(objs, res):
        o.solve_async(st, it, args["mu0"], args["maxIter"])
    for o, (st, it) in zip(objs, res):
        y = o.getOutputs()                      # synchronises the handle
        assert np.array_equal(st, st0) and np.array_equal(it, it0)
        for a, b in zip(y, y0):
            assert np.array_equal(a, b)
        o.close()


def test_queued_batches_on_one_handle_with_async_outputs():
    """set + tc_solve_async + tc_get_outputs_async twice on ONE handle without a host synchronisation in between
    (two sets of pinned host buffers): both batches come back as the blocking calls deliver them"""
    import torch
    gen, args = generated("mpc_dcmotor")
    batch = 900
    ins = [problems.synthetic_inputs("mpc_dcmotor", gen, batch, seed=s) for s in (21, 22)]
    want = []
    for P, V in ins:
        ref = _solve(gen, args, P, V, batch)
        st, it, _ = ref.solve(args["mu0"], args["maxIter"], -1)
        want.append((st, it, ref.getOutputs()))
        ref.close()
    sizes = [int(np.prod(s)) if s else 1 for _, s in gen.outputs]
    obj = gen(batch)
    bufs = [([torch.empty((batch, n), dtype=torch.float64).pin_memory() for n in sizes],
             torch.empty(batch, dtype=torch.int32).pin_memory(), torch.empty(batch, dtype=torch.int32).pin_memory())
            for _ in ins]
    for (P, V), (outs, st, it) in zip(ins, bufs):
        for n, _ in gen.parameters:
            getattr(obj, "setP_" + n)(P[n])
        for n, _ in gen.variables:
            getattr(obj, "setV_" + n)(V[n])
        obj.solve_async(st.numpy(), it.numpy(), args["mu0"], args["maxIter"])
        obj.get_outputs_ptrs([t.data_ptr() for t in outs], device=False, sync=False)
    obj.synchronize()
    for (st0, it0, y0), (outs, st, it) in zip(want, bufs):
        assert np.array_equal(st.numpy(), st0) and np.array_equal(it.numpy(), it0)
        ys = list(y0.values()) if isinstance(y0, dict) else list(y0)
        for a, b, (_, shape) in zip(outs, ys, gen.outputs):        # same unpacking as BatchedSolver.getOutputs
            assert np.array_equal(a.numpy().T.reshape(tuple(shape) + (batch,), order="F"), b)
    obj.close()


def test_tiles_option_is_bit_identical():
    """two instances per lane (emitter option tiles=2) walk the same code: identical results"""
    gen, args = generated("sls")
    batch = 1000                                # not a multiple of the CTA size: ragged tail
    P, V = problems.synthetic_inputs("sls", gen, batch, seed=6)
    outs = []
    for opts in ({}, dict(tiles=2)):
        obj = _solve(gen, args, P, V, batch, **opts)
        st, it, _ = obj.solve(args["mu0"], args["maxIter"], -1)
        outs.append((st, it, obj.getOutputs()))
        obj.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    for a, b in zip(outs[0][2], outs[1][2]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("opts", [dict(pair=1, max_loads=14), dict(scache=24), dict(pair=1, max_loads=14, scache=12),
                                  dict(lockstep=6, warps_per_sm=24), dict(prefetch=0), dict(wide=1), dict(roll=20)],
                         ids=["pair", "scache", "pair+scache", "lockstep6", "noprefetch", "wide", "roll"])
@pytest.mark.parametrize("name,batch", [("sls", 1000), ("mpc_dcmotor", 700)])
def test_layout_options_are_bit_identical(name, batch, opts):
    """emitter options that change where values live, not what is computed: `pair` (two instances per lane in
    interleaved columns, 128-bit loads and stores; the set/gather/compaction kernels follow the layout through
    tc_device_image::ilv), `wide` (tiles of 64 instances: two lockstep warps share 512-byte segments), `scache` (per-lane software cache in shared memory, offline Belady plan), CTAs of 6
    lockstep warps, no L2 prefetch, `roll` (strided repeats of one instruction as device loops: the reference's
    loop rolling) -- same status, iterations and outputs as the default module, on a ragged batch,
    and the same again after a set + re-solve on the same handle (stale lazy-evaluation state of the partner)"""
    gen, args = generated(name)
    P, V = problems.synthetic_inputs(name, gen, batch, seed=11)
    P2, V2 = problems.synthetic_inputs(name, gen, batch, seed=12)
    res = []
    for o in ({}, opts):
        obj = _solve(gen, args, P, V, batch, **o)
        st, it, _ = obj.solve(args["mu0"], args["maxIter"], -1)
        y = obj.getOutputs()
        for n, _ in gen.parameters:                 # second solve on the same handle: half of the instances get
            v = np.array(P[n], copy=True)           # new parameters, all get new initial values
            v[..., ::2] = P2[n][..., ::2]
            getattr(obj, "setP_" + n)(v)
        for n, _ in gen.variables:
            getattr(obj, "setV_" + n)(V2[n])
        st2, it2, _ = obj.solve(args["mu0"], args["maxIter"], -1)
        res.append((st, it, y, st2, it2, obj.getOutputs()))
        obj.close()
    a, b = res
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
    for x, y in zip(list(a[2]) + list(a[5]), list(b[2]) + list(b[5])):
        assert np.array_equal(x, y)
    assert (a[0] == 0).mean() > 0.5


@pytest.mark.parametrize("backend", ["compiled", "tape"])
def test_allow_save_keeps_the_kkt_values(backend):
    """allowSave / saveIter (ipmPD_CSsolver.c:136,330-332,561-566,671-676,877-882): a solver generated with
    allowSave keeps, per instance, the KKT matrix values of iteration saveIter (saveIter = 0: at the successful
    exit) and returns them as the last output -- identical to what the reference's driver hands to saveWW__, with
    the solve itself unchanged; phased (compiled) and single-launch (tape) execution"""
    batch = 300
    gen0, args = generated("mpc_dcmotor")
    gen, _ = generated("mpc_dcmotor", allowSave=True)
    P, V = problems.synthetic_inputs("mpc_dcmotor", gen0, batch, seed=6)
    for save_iter in (3, 0):
        st_o, it_o, out_o, _ = _oracle(gen, "mpc_dcmotor").solve_batch(P, V, args["mu0"], args["maxIter"], saveIter=save_iter)
        assert (st_o == 0).all() and it_o.min() > 3
        obj = _solve(gen, args, P, V, batch, backend=backend)
        st, it, _ = obj.solve(args["mu0"], args["maxIter"], save_iter)
        outs = obj.getOutputs()
        outs = outs if isinstance(outs, dict) else {n: o for (n, _), o in zip(gen.outputs, outs)}
        obj.close()
        assert np.array_equal(st, st_o) and np.array_equal(it, it_o)
        for k in out_o:
            a, b = np.asarray(outs[k]), np.asarray(out_o[k])
            assert np.abs(a - b).max() <= 1e-8 * max(1.0, np.abs(b).max()), (save_iter, k)
        assert np.abs(outs["WW_values__"]).sum() > 0


def test_componentwise_user_function():
    """I_componentwise (user C text, instructionsTableUTHash.c:1333-1345; Tcalculus.m:1836-1880): isa_probe with its
    sine term written as `componentwise(x,sin,'sin(%s)',d1,d2)` -- compiled into the module as CUDA C -- against
    the reference's emitter and driver, and bit-identical to the module that uses the built-in opcodes"""
    batch = 512
    gen0, args = generated("isa_probe")
    gen1, _ = generated("isa_probe", cw=True)
    P, V = problems.synthetic_inputs("isa_probe", gen0, batch, seed=4)
    st_o, it_o, out_o, _ = _oracle(gen1, "isa_probe").solve_batch(P, V, args["mu0"], args["maxIter"])
    res = []
    for gen in (gen0, gen1):
        obj = _solve(gen, args, P, V, batch)
        st, it, _ = obj.solve(args["mu0"], args["maxIter"], -1)
        outs = obj.getOutputs()
        res.append((st, it, list(outs.values()) if isinstance(outs, dict) else list(outs)))
        obj.close()
    st, it, outs = res[1]
    assert np.array_equal(st, st_o) and np.abs(it.astype(int) - it_o.astype(int)).max() <= 1
    assert (st_o == 0).mean() > 0.5
    for (n, _), y in zip(gen1.outputs, outs):
        a, b = np.asarray(y)[..., st_o == 0], np.asarray(out_o[n])[..., st_o == 0]
        assert np.abs(a - b).max() <= 1e-8 * max(1.0, np.abs(b).max()), n
    assert np.array_equal(res[0][0], st) and np.array_equal(res[0][1], it)
    for a, b in zip(res[0][2], outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("driver", ["lanes", "queue"])
@pytest.mark.parametrize("name,batch,kw", [("sls", 5000, {}), ("mpc_quadcopter", 256, {"T_": 20, "useInertia": True}),
                                           ("sls", 3000, AFFINE)])
def test_lane_decoupled_driver_parity(name, batch, kw, driver):
    """emitter option driver="lanes" (csrc/tc_driver_lanes.cuh: every lane runs the reference's loop as a
    coroutine) and driver="queue" (the same, with resident warps pulling instances from a work queue into
    a work scratchbook): same status, same iteration counts, same outputs as the CPU oracle"""
    gen, args = generated(name, **kw)
    P, V = problems.synthetic_inputs(name, gen, batch, seed=2)
    st_o, it_o, out_o, _ = _oracle(gen, name).solve_batch(P, V, args["mu0"], args["maxIter"])
    obj = _solve(gen, args, P, V, batch, driver=driver)
    st, it, _ = obj.solve(args["mu0"], args["maxIter"], -1)
    outs = obj.getOutputs()
    if not isinstance(outs, dict):
        outs = {n: o for (n, _), o in zip(gen.outputs, outs)}
    obj.close()
    assert np.array_equal(st, st_o) and np.array_equal(it, it_o)
    ok = st_o == 0
    for k in out_o:
        a, b = outs[k][..., ok], out_o[k][..., ok]
        scale = np.maximum(np.abs(b).max(axis=tuple(range(b.ndim - 1)), keepdims=True), 1e-300)
        assert np.max(np.abs(a - b) / scale) <= 1e-8, k


def _compare_with_oracle(gen, name, P, V, mu0, maxIter, batch, exact_iter=True, **opts):
    st_o, it_o, out_o, _ = _oracle(gen, name).solve_batch(P, V, mu0, maxIter)
    obj = _solve(gen, None, P, V, batch, **opts)
    st, it, _ = obj.solve(mu0, maxIter, -1)
    obj.close()
    assert np.array_equal(st, st_o), f"status: gpu {np.unique(st)} oracle {np.unique(st_o)}"
    if exact_iter:
        assert np.array_equal(it, it_o)
    else:
        assert np.abs(it.astype(int) - it_o.astype(int)).max() <= 1
    return st, it


@pytest.mark.parametrize("driver", ["warp", "lanes"])
def test_forced_exit_codes_match_the_reference_driver(driver):
    """every exit path of ipmPD_CSsolver.c:308-357,696-705,885-920 is produced on purpose and compared with the
    reference driver: status 8 with its diagnostic bits (16 gradient, 32 equality, 64 gap, 128 mu, 1024/1536/1792
    step sizes, 2048 regulariser) through a small maxIter, status 1 (infeasible start), status 2 (negative dual
    through mu0<0) and status 4 (NaN parameter)"""
    seen = set()
    for name, kw, batch in (("sls", {}, 256), ("mpc_dcmotor", {}, 256), ("mpc_quadcopter", {"T_": 20, "useInertia": True}, 128)):
        gen, args = generated(name, **kw)
        P, V = problems.synthetic_inputs(name, gen, batch, seed=11)
        for maxIter in (1, 2, 3, 5, 8):                       # status 8 | diagnostic bits
            st, _ = _compare_with_oracle(gen, name, P, V, args["mu0"], maxIter, batch, driver=driver)
            seen.update(int(x) for x in np.unique(st))
        st, _ = _compare_with_oracle(gen, name, P, V, -args["mu0"], args["maxIter"], batch, driver=driver)   # lambda=mu/F<0
        seen.update(int(x) for x in np.unique(st))
    # status 1: start outside the box; status 4: NaN in the data
    gen, args = generated("sls")
    batch = 256
    P, V = problems.synthetic_inputs("sls", gen, batch, seed=12)
    V["x"][:, ::2] = .2                                      # violates x <= .05 on every second instance
    P["b"][0, 1::4] = np.nan
    st, _ = _compare_with_oracle(gen, "sls", P, V, args["mu0"], args["maxIter"], batch, driver=driver)
    seen.update(int(x) for x in np.unique(st))
    assert {1, 2, 4} <= seen, seen
    assert any(s & 8 for s in seen) and any(s & 16 for s in seen) and any(s & 64 for s in seen) and \
        any(s & 128 for s in seen), seen
    print("exit codes produced and matched:", sorted(seen))


def test_queue_driver_set_then_get_outputs_without_solve():
    """after a work-queue solve the optimum is back in the batch scratchbook: changing a parameter and calling
    getOutputs (no new solve) re-evaluates the outputs at the optimum, like the other drivers and the reference"""
    gen, args = generated("sls")
    batch = 3000
    P, V = problems.synthetic_inputs("sls", gen, batch, seed=13)
    res = []
    for driver in ("warp", "queue"):
        obj = _solve(gen, args, P, V, batch, driver=driver)
        obj.solve(args["mu0"], args["maxIter"], -1)
        obj.setP_b(P["b"] * 1.5)
        res.append(obj.getOutputs())
        obj.close()
    for a, b in zip(*res):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name,batch,kw,phases", [("sls", 5000, {}, ()), ("sls", 33, {}, (1,)), ("mpc_dcmotor", 4097, {}, (3, 1, 2)),
                                                  ("mpc_quadcopter", 1000, {"T_": 20, "useInertia": True}, ()),
                                                  ("mpc_dcmotor", 2000, AFFINE, (2,))])
def test_phased_solve_is_bit_identical_to_the_single_launch_solve(name, batch, kw, phases):
    """the phased solve (phases of a few Newton iterations, unfinished instances compacted into full warps in
    between, everything moved home at the end: csrc/tc_driver.cuh, tc_abi.cu) gives exactly the status, iteration
    counts and outputs of the single-launch solve -- also after a set_<parameter> + getOutputs without a new solve"""
    gen, args = generated(name, **kw)
    P, V = problems.synthetic_inputs(name, gen, batch, seed=21)
    res = []
    for phased in (False, True):
        obj = _solve(gen, args, P, V, batch)
        obj.set_phased(phased, phases)
        st, it, _ = obj.solve(args["mu0"], args["maxIter"], -1)
        y = obj.getOutputs()
        y = list(y.values()) if isinstance(y, dict) else y
        first = gen.parameters[0][0]
        getattr(obj, "setP_" + first)(np.asarray(P[first]) * 1.25)       # outputs re-evaluated at the optimum
        y2 = obj.getOutputs()
        y2 = list(y2.values()) if isinstance(y2, dict) else y2
        st2, it2, _ = obj.solve(args["mu0"], args["maxIter"], -1)         # and a second solve on the same handle
        res.append((st, it, y, y2, st2, it2))
        obj.close()
    a, b = res
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5])
    for u, v in list(zip(a[2], b[2])) + list(zip(a[3], b[3])):
        assert np.array_equal(u, v, equal_nan=True)
    assert len(np.unique(a[1])) > 1          # instances really finish at different iterations


def test_receding_horizon_warm_start_stays_on_device():
    """The MPC loop of examples/mpcmhe/mpc_dcmotor.m:192-225 -- solve, apply u(:,1) to the plant, shift the
    reference, warm-start the next solve with out.xWarm/out.uWarm -- with the warm start fed back ON THE DEVICE
    (tc_feedback) against the oracle run through host arrays like the reference script: identical status,
    iterations and outputs at every step of the closed loop."""
    from oracle.oracle import Oracle
    name, B, steps = "mpc_dcmotor", 256, 6
    gen, args = generated(name)
    P, V = problems.synthetic_inputs(name, gen, B, seed=11)
    P = {k: np.array(v, dtype=float) for k, v in P.items()}
    Ts, p, k = .1, -2., 1.
    # exact zero-order-hold discretisation of xdot=[0 1;0 p]x+[0;k]u
    e = np.exp(p * Ts)
    Ad = np.array([[1, (e - 1) / p], [0, e]])
    Bd = np.array([k * ((e - 1) / p - Ts) / p, k * (e - 1) / p])
    T_ = P["ref"].shape[1]
    t0 = np.random.default_rng(3).uniform(0, 4 * np.pi, B)
    oracle = Oracle(gen, "ref")
    obj = gen(B)
    for n, _ in gen.parameters:
        getattr(obj, "setP_" + n)(P[n])
    for n, _ in gen.variables:
        getattr(obj, "setV_" + n)(V[n])
    Vk = dict(V)
    warm_iters, cold_iters = 0, 0
    for step in range(steps):
        st, it, _ = obj.solve(args["mu0"], args["maxIter"], -1)
        out = obj.getOutputs()
        st_o, it_o, out_o, _ = oracle.solve_batch(P, Vk, args["mu0"], args["maxIter"])
        assert np.array_equal(st, st_o), step
        assert np.array_equal(it, it_o), step
        for key in out_o:
            assert np.array_equal(out[key], out_o[key]), (step, key)
        cold_iters += int(it.sum()) if step == 0 else 0
        warm_iters += int(it.sum()) if step == steps - 1 else 0
        # device: warm start without a host round trip; host oracle: the reference script's arrays
        obj.feedback({"xWarm": "x", "uWarm": "u"})
        Vk = dict(x=out_o["xWarm"], u=out_o["uWarm"])
        xinit = P["xinit"][:, 0, :]
        xinit = Ad @ xinit + Bd[:, None] * out_o["u"][0, 0, :][None, :]
        tt = t0[None, :] + Ts * (step + 1 + np.arange(T_))[:, None]
        P["xinit"] = xinit[:, None, :]
        P["ref"] = (-.35 * np.sign(np.sin(.5 * tt)))[None, :, :]
        obj.setP_xinit(P["xinit"])
        obj.setP_ref(P["ref"])
    obj.close()
    print(f"closed loop {steps} steps x {B} instances: {cold_iters / B:.1f} iterations cold, {warm_iters / B:.1f} warm")


def test_feedback_rejects_mismatched_sizes():
    gen, _ = generated("mpc_dcmotor")
    obj = gen(64)
    with pytest.raises(Exception):
        obj.feedback({"J": "x"})
    obj.close()
