"""Lowering of a compiled problem to the device image executed by the sm_100a kernels.

This replaces the reference's C emitter (``writeCinstructionsC``,
``lib/csparse/instructionsTableUTHash.c:790-1609`` + the wrappers of
``writeCfunctionpergroup.m:513-852``): instead of C text compiled per problem, the
straight-line code of every dependency group becomes a compact *tape* of 32-bit words
that the generic kernels in ``csrc/`` interpret with one instance per lane, so the same
pre-built ``.so`` serves any problem structure.

Tape, per instruction::

    word 0   opcode (bits 0-7, reference opcode numbers) | n_operands << 8
    word 1   result slot
    [word 2  m = factors per product, I_sumprod only]
    ...      operand words: bits 0-29 scratch slot or constant-pool index,
                            bit 30 = operand is a constant, bit 31 = negate (I_sum)

The evaluation order inside an instruction is the C expression order of the reference
emitter, and the kernels use separate multiply/add roundings,
so results are bit-identical to the CPU oracle wherever no libm function is involved.

Image = 64 int64 header + 16 double header + sections (see ``serialize``).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from . import itable as it
from .csparse import CompiledProblem
from .optimize import GeneratedSolver

MAGIC = 0x3030324243540001  # "\x01\0TCB200" little-endian tag + format version
CONST_BIT = 1 << 30
NEG_BIT = 1 << 31
NO_SLOT = 0xFFFFFFFF

# driver function ids -- MUST match enum tc_fn in csrc/tc_image.h
FN = {name: k for k, name in enumerate([
    "initPrimal__", "scaleIneq__", "scaleCost__", "getScale4Cost__", "setMu__", "initDualIneq__",
    "initDualEqX__", "setAddEye2HessianU__", "setAddEye2HessianEq__", "getNorminf_Grad__",
    "getNorminf_G__", "getGapMinFMinLambda__", "getCurvature__", "getDirectionError__",
    "getHessInertia__", "getMaxAlphas_s__", "setAlphaPrimal__", "getMinF_s__", "setAlphaDualEq__",
    "setAlphaDualIneq__", "updatePrimalDual__", "getJ__", "getOutputs"])}
FN_USER0 = len(FN)
KIND_NONE, KIND_SET, KIND_GET, KIND_COPY = 0, 1, 2, 3


def _mask128(groups) -> Tuple[int, int]:
    lo = hi = 0
    for g in groups:
        if g >= 128:
            raise ValueError("more than 128 dependency groups")
        if g < 64:
            lo |= 1 << g
        else:
            hi |= 1 << (g - 64)
    return lo, hi


def lower(gen: GeneratedSolver) -> Dict[str, np.ndarray]:
    pr: CompiledProblem = gen.problem
    tab, slot = pr.tab, pr.slot
    if pr.n_slots >= CONST_BIT:
        raise ValueError("scratchbook too large for the 30-bit slot field")
    consts: List[float] = []
    cindex: Dict[float, int] = {}

    def cword(v: float) -> int:
        k = cindex.get(v)
        if k is None:
            k = cindex[v] = len(consts)
            consts.append(v)
        return CONST_BIT | k

    def opword(i: int) -> int:
        t = tab.types[i]
        if t == it.I_load:
            return cword(tab.params[i][0])
        if t == it.I_2:
            return cword(2.0)
        return int(slot[i])

    tape: List[int] = []
    group_start = [0]
    operand_reads = []          # per group: scratch operand loads (constants excluded)
    for members in pr.groups:
        reads = 0
        for i in members:
            t, p, ops = tab.types[i], tab.params[i], tab.operands[i]
            if t in (it.I_load, it.I_2):
                words = [cword(p[0] if t == it.I_load else 2.0)]
                tape += [it.I_sum | (1 << 8), int(slot[i])] + words
                continue
            words = [opword(o) for o in ops]
            if t == it.I_sum:
                words = [w | (NEG_BIT if s < 0 else 0) for w, s in zip(words, p)]
            tape += [t | (len(words) << 8), int(slot[i])]
            if t == it.I_sumprod:
                tape.append(int(p[0]))
            tape += words
            reads += sum(1 for w in words if not w & CONST_BIT)
        group_start.append(len(tape))
        operand_reads.append(reads)

    # ---- function table ------------------------------------------------------------------------
    user_sets = [s for s in pr.sets.values() if s.kind in ("parameter", "variable")]
    n_funcs = FN_USER0 + len(user_sets)
    funcs = np.zeros((n_funcs, 6), dtype=np.uint64)   # parents lo/hi, children lo/hi, io_off|io_cnt<<32, kind
    io: List[int] = []

    def put(fid: int, kind: int, parents, children, words: List[int]) -> None:
        plo, phi = _mask128(parents)
        clo, chi = _mask128(children)
        funcs[fid] = (plo, phi, clo, chi, len(io) | (len(words) << 32), kind)
        io.extend(words)

    def src_word(x) -> int:
        return NO_SLOT if x is None else opword(x)

    for name, fid in FN.items():
        if name in pr.sets:
            s = pr.sets[name]
            put(fid, KIND_SET, [], s.children, [int(slot[d]) for d in s.dest])
        elif name in pr.gets:
            g = pr.gets[name]
            put(fid, KIND_GET, g.parents, [], [src_word(x) for src in g.sources for x in src])
        elif name in pr.copies:
            c = pr.copies[name]
            words = []
            for d, s in zip(c.dest, c.src):
                if d != s:
                    words += [int(slot[d]), src_word(s)]
            put(fid, KIND_COPY, c.parents, c.children, words)
    # user sets in API order: parameters, then optimisation variables
    set_order = [f"setP_{n}" for n, _ in gen.parameters] + [f"setV_{n}" for n, _ in gen.variables]
    set_fids = []
    for k, name in enumerate(set_order):
        s = pr.sets[name]
        put(FN_USER0 + k, KIND_SET, [], s.children, [int(slot[d]) for d in s.dest])
        set_fids.append(FN_USER0 + k)
    out_sizes = [int(np.prod(s)) if s else 1 for _, s in gen.outputs]

    return dict(tape=np.asarray(tape, dtype=np.uint32), group_start=np.asarray(group_start, dtype=np.uint32),
                consts=np.asarray(consts if consts else [0.0], dtype=np.float64), funcs=funcs,
                io=np.asarray(io if io else [0], dtype=np.uint32),
                set_fids=np.asarray(set_fids, dtype=np.uint32), out_sizes=np.asarray(out_sizes, dtype=np.uint32),
                operand_reads=np.asarray(operand_reads, dtype=np.int64))


def serialize(gen: GeneratedSolver) -> bytes:
    L = lower(gen)
    d = gen.defines
    ih = np.zeros(64, dtype=np.int64)
    ih[0] = MAGIC
    ih[1:9] = (gen.problem.n_slots, len(gen.problem.groups), len(L["funcs"]), len(L["set_fids"]),
               len(L["out_sizes"]), len(L["tape"]), len(L["consts"]), len(L["io"]))
    ih[16:19] = (d["nU"], d["nG"], d["nF"])
    ih[19:27] = (d["scaleCost"], d["scaleInequalities"], d["coupledAlphas"], d["setAddEye2Hessian"],
                 d["adjustAddEye2Hessian"], d["useInertia"], d["skipAffine"], d["delta"])
    dh = np.zeros(16, dtype=np.float64)
    dh[0:8] = (d["gradTolerance"], d["addEye2HessianUtolerance"], d["equalTolerance"], d["desiredDualityGap"],
               d["alphaMin"], d["alphaMax"], d["muFactorAggressive"], d["muFactorConservative"])

    def pad8(b: bytes) -> bytes:
        return b + b"\0" * (-len(b) % 8)

    parts = [ih.tobytes(), dh.tobytes()]
    for key in ("group_start", "tape", "consts", "funcs", "io", "set_fids", "out_sizes"):
        parts.append(pad8(np.ascontiguousarray(L[key]).tobytes()))
    return b"".join(parts)
