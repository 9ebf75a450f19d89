"""Computation-graph files (.cg/.cgc/.cgio/.cgs): the reference's wire format for the scalar instruction
graph (lib/@csparse/saveScalarized.m:108-273, lib/CGregistration.m:65-133, doc/computationgraphs.tex:84-205).
MATLAB is not available here, so the files under test are written by tenscalc_b200.cgraph.save_scalarized,
which follows saveScalarized.m record for record (including its 1-based get/set variables in .cgio)."""
import os
import struct

import numpy as np
import pytest

from conftest import generated
from tenscalc_b200 import cgraph, problems
from tenscalc_b200.csparse import fingerprint


@pytest.mark.parametrize("name", ["sls", "mpc_dcmotor", "l1l2"])
def test_round_trip_reproduces_the_problem(tmp_path, name):
    gen, _ = generated(name)
    base = str(tmp_path / name)
    stats = cgraph.save_scalarized(gen, base)
    assert stats["ios"] == len(gen.problem.gets) + len(gen.problem.sets) + len(gen.problem.copies)
    opts = {k: gen.defines[k] for k in ("gradTolerance", "equalTolerance", "desiredDualityGap", "alphaMin", "alphaMax",
                                        "muFactorAggressive", "muFactorConservative", "addEye2HessianUtolerance")}
    g2 = cgraph.load_scalarized(base, output_shapes={"getOutputs": [s for _, s in gen.outputs]}, **opts)
    pr, p2 = gen.problem, g2.problem
    assert p2.tab.types == pr.tab.types and p2.tab.operands == pr.tab.operands and p2.tab.params == pr.tab.params
    assert p2.groups == pr.groups and np.array_equal(p2.slot, pr.slot)
    assert p2.defines == pr.defines                       # nU/nG/nF and the option switches are recovered from the graph
    assert fingerprint(p2) == fingerprint(pr)             # => same emitted CUDA, same oracle library
    assert g2.classname == gen.classname
    assert [n for n, _ in g2.parameters] == [n for n, _ in gen.parameters]
    assert [n for n, _ in g2.variables] == [n for n, _ in gen.variables]
    assert [int(np.prod(s)) if s else 1 for _, s in g2.parameters] == [int(np.prod(s)) if s else 1 for _, s in gen.parameters]


def test_file_layout_follows_saveScalarized(tmp_path):
    """one magic byte 32, int32 little-endian records; .cg: [type,nIn,nOut,in..,out..] / [0,const#,out];
    I_sum carries its +-1 parameter vector as a leading constant input (saveScalarized.m:117-120,146,166-172)"""
    gen, _ = generated("sls")
    base = str(tmp_path / "sls")
    cgraph.save_scalarized(gen, base)
    for ext in ("cg", "cgc", "cgio", "cgs"):
        assert open(f"{base}.{ext}", "rb").read(1) == bytes([32])
    constants, nodes, ios, names = cgraph.read_files(base)
    tab = gen.problem.tab
    by_out = {n[2][0]: n for n in nodes if n[0] not in (cgraph.CG_FT2S,)}
    some_sum = next(i for i in range(len(tab)) if tab.types[i] == 3)            # I_sum
    t, ins, outs = by_out[some_sum]
    assert t == 20001 and outs == [some_sum] and tuple(ins[1:]) == tab.operands[some_sum]
    par = constants[by_out[ins[0]][1][0]]
    assert par.dtype == np.int64 and tuple(float(x) for x in par.reshape(-1)) == tab.params[some_sum]
    kinds = [k for k, _, _ in ios]
    assert kinds == sorted(kinds, key=lambda k: {2: 0, 1: 1, 3: 2}[k])          # gets, then sets, then copies (:188-250)
    assert f"{gen.classname}_getOutputs" in names.values() and "updatePrimalDual__" in names.values()
    # the reference writes get/set variables 1-based (:199,219): the reader detects it ...
    g2 = cgraph.load_scalarized(base)
    assert g2.info["io_variable_base"] == 1
    # ... and also accepts the 0-based convention of doc/computationgraphs.tex
    raw = bytearray(open(base + ".cgio", "rb").read())
    pos, out = 1, bytearray(raw[:1])
    while pos < len(raw):
        kind, n = struct.unpack_from("<ii", raw, pos)
        cnt = n * (2 if kind == 3 else 1)
        vals = list(struct.unpack_from(f"<{cnt}i", raw, pos + 8))
        if kind != 3:
            vals = [v - 1 for v in vals]
        out += struct.pack(f"<ii{cnt}i", kind, n, *vals)
        pos += 8 + 4 * cnt
    open(base + ".cgio", "wb").write(out)
    g3 = cgraph.load_scalarized(base)
    assert g3.info["io_variable_base"] == 0 and fingerprint(g3.problem) == fingerprint(g2.problem)


def test_unregistered_types_and_broken_files(tmp_path):
    gen, _ = generated("mpc_quadcopter", T_=20, useInertia=True)     # uses I_heaviside: not in CGregistration.m
    with pytest.raises(cgraph.CGFormatError, match="missing type"):
        cgraph.save_scalarized(gen, str(tmp_path / "strict"), extended=False)          # saveScalarized.m:183
    base = str(tmp_path / "q")
    cgraph.save_scalarized(gen, base)
    g2 = cgraph.load_scalarized(base, output_shapes={"getOutputs": [s for _, s in gen.outputs]})
    assert fingerprint(g2.problem) == fingerprint(gen.problem) and g2.defines["useInertia"] == 1
    with pytest.raises(ValueError):
        cgraph.save_scalarized(gen, str(tmp_path / "name.cg"))                          # :113-115
    data = open(base + ".cg", "rb").read()
    open(base + ".cg", "wb").write(data[:len(data) // 2 + 1])                          # truncated mid-record
    with pytest.raises(cgraph.CGFormatError):
        cgraph.load_scalarized(base)
    open(base + ".cg", "wb").write(b"\x07" + data[1:])
    with pytest.raises(cgraph.CGFormatError, match="magic"):
        cgraph.load_scalarized(base)


def test_loaded_problem_solves_like_the_generated_one(tmp_path):
    """the solver rebuilt from the files runs through the CPU oracle with identical results"""
    from oracle.oracle import Oracle
    gen, args = generated("mpc_dcmotor")
    base = str(tmp_path / "m")
    cgraph.save_scalarized(gen, base)
    g2 = cgraph.load_scalarized(base, output_shapes={"getOutputs": [s for _, s in gen.outputs]})
    P, V = problems.synthetic_inputs("mpc_dcmotor", gen, 8, seed=3)
    st, it, out, _ = Oracle(gen, "port").solve_batch(P, V, args["mu0"], args["maxIter"])
    P2 = {n: np.asarray(P[n]).reshape(tuple(s) + (8,), order="F") for n, s in g2.parameters}    # the files carry no shapes: flat
    V2 = {n: np.asarray(V[n]).reshape(tuple(s) + (8,), order="F") for n, s in g2.variables}
    st2, it2, out2, _ = Oracle(g2, "port").solve_batch(P2, V2, args["mu0"], args["maxIter"])
    assert np.array_equal(st, st2) and np.array_equal(it, it2)
    for (n1, _), (n2, _) in zip(gen.outputs, g2.outputs):
        assert np.array_equal(np.asarray(out[n1]).reshape((-1, 8), order="F"), np.asarray(out2[n2]).reshape((-1, 8), order="F"))


@pytest.mark.gpu
def test_loaded_problem_on_the_gpu(tmp_path):
    gen, args = generated("sls")
    base = str(tmp_path / "s")
    cgraph.save_scalarized(gen, base)
    g2 = cgraph.load_scalarized(base, output_shapes={"getOutputs": [s for _, s in gen.outputs]})
    B = 512
    P, V = problems.synthetic_inputs("sls", gen, B, seed=4)
    res = []
    for g, Pm, Vm in ((gen, P, V), (g2, {"A": P["A"].reshape(-1, B, order="F"), "b": P["b"]}, V)):
        obj = g(B)
        for n, _ in g.parameters:
            getattr(obj, "setP_" + n)(Pm[n])
        obj.setV_x(Vm["x"])
        st, it, _ = obj.solve(args["mu0"], args["maxIter"], -1)
        res.append((st, it, [np.asarray(o).reshape(-1, B) for o in obj.getOutputs()]))
        obj.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    for a, b in zip(res[0][2], res[1][2]):
        assert np.array_equal(a, b)


def test_componentwise_text_survives_the_files(tmp_path):
    """I_componentwise has no registered code in saveScalarized.m (:138-195): the extended writer stores the C text as
    a parameter vector (like the sign vector of a sum) under a private type code, the reader restores it -- same
    instruction table, same fingerprint, hence the same compiled module; the strict writer refuses"""
    from tenscalc_b200 import itable as it
    gen, _ = generated("isa_probe", cw=True)
    base = str(tmp_path / "cw")
    cgraph.save_scalarized(gen, base)
    g2 = cgraph.load_scalarized(base, output_shapes={"getOutputs": [s for _, s in gen.outputs]})
    pr, p2 = gen.problem, g2.problem
    assert p2.tab.types == pr.tab.types and p2.tab.operands == pr.tab.operands and p2.tab.params == pr.tab.params
    texts = sorted(it.componentwise_text(p2.tab.params[i], "%s") for i in range(len(p2.tab))
                   if p2.tab.types[i] == it.I_componentwise)
    assert texts == ["cos(%s)", "sin(%s)"]
    assert fingerprint(p2) == fingerprint(pr)
    with pytest.raises(cgraph.CGFormatError):
        cgraph.save_scalarized(gen, str(tmp_path / "strict"), extended=False)
