"""Multi-rank host logic on CPU: world_size-2 (and 3) gloo processes, each owning a contiguous
range of the batch; per-rank solves are done by the CPU oracle standing in for the GPU solver
(same setP_/setV_/solve/getOutputs surface), then gathered and compared with a single-process
solve of the whole batch.  No collective is used on the data path."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import generated
from tenscalc_b200 import problems
from tenscalc_b200.sharding import shard_range


def test_shard_range_covers_batch_exactly():
    for batch in (1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(batch, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


class _OracleBackedSolver:
    """stand-in with the generated class's surface, for CPU tests of the sharding logic only"""

    def __init__(self, gen, batch, device=None):
        from oracle.oracle import Oracle
        self.gen, self.batch = gen, batch
        self.oracle = Oracle(gen, "port")
        self.P, self.V = {}, {}

    def __getattr__(self, name):
        if name.startswith("setP_"):
            return lambda v: self.P.__setitem__(name[5:], np.asarray(v))
        if name.startswith("setV_"):
            return lambda v: self.V.__setitem__(name[5:], np.asarray(v))
        raise AttributeError(name)

    def solve(self, mu0, maxIter, saveIter=-1, addEye2Hessian=(1e-9, 1e-9)):
        bc = lambda d, spec: {n: (v if v.shape != tuple(s) else np.repeat(v[..., None], self.batch, -1))
                              for (n, s) in spec for v in [d[n]]}
        st, it, self.out, sec = self.oracle.solve_batch(bc(self.P, self.gen.parameters), bc(self.V, self.gen.variables),
                                                        mu0, maxIter, saveIter, addEye2Hessian)
        return st, it, float(sec.sum())

    def getOutputs(self):
        return self.out if self.gen.output_is_struct else [self.out[n] for n, _ in self.gen.outputs]

    def close(self):
        pass


def _worker(rank, world, port, name, batch, seed, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tenscalc_b200.sharding import ShardedSolver
    gen, args = problems.BUILDERS[name]()
    P, V = problems.synthetic_inputs(name, gen, batch, seed=seed)
    s = ShardedSolver(gen, batch, rank, world, solver_factory=lambda g, b, d: _OracleBackedSolver(g, b, d))
    for n, _ in gen.parameters:
        getattr(s, "setP_" + n)(P[n])
    for n, _ in gen.variables:
        getattr(s, "setV_" + n)(V[n])
    st, it, _ = s.solve(args["mu0"], args["maxIter"])
    res = s.gather(st, it, s.getOutputs())
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("name,batch,world", [("sls", 37, 2), ("mpc_dcmotor", 10, 3)])
def test_sharded_solve_equals_single_process(name, batch, world):
    from oracle.oracle import Oracle
    gen, args = generated(name)
    Oracle(gen, "port").build()
    seed = 9
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, batch, seed, q))
             for port in [_free_port()] for r in range(world)]
    for p in procs:
        p.start()
    st, it, outs = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    P, V = problems.synthetic_inputs(name, gen, batch, seed=seed)
    st1, it1, out1, _ = Oracle(gen, "port").solve_batch(P, V, args["mu0"], args["maxIter"])
    assert np.array_equal(st, st1) and np.array_equal(it, it1)
    outs = outs if isinstance(outs, dict) else dict(zip([n for n, _ in gen.outputs], outs))
    for k, v in out1.items():
        assert np.array_equal(outs[k], v), k
