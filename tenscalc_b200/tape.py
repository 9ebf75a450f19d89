"""Lowering of a compiled problem to the device image executed by the sm_100a kernels.

This replaces the reference's C emitter (``writeCinstructionsC``,
``lib/csparse/instructionsTableUTHash.c:790-1609`` + the wrappers of
``writeCfunctionpergroup.m:513-852``): instead of C text compiled per problem, the
straight-line code of every dependency group becomes a compact *tape* that the generic
kernels in ``csrc/`` interpret with one instance per lane, so the same pre-built ``.so``
serves any problem structure.

Tape = array of 128-bit CELLS.  A cell is one step of an accumulator machine with two
registers per lane (``acc``, ``p``) and two memory operands::

    word 0   op (bits 0-6) | flags (bits 7-15) | ring slot of the result (bits 16-23)
    word 1   operand A   early operand (scratchbook): byte offset slot*256 inside the warp's tile
    word 2   operand B   late operand (F_ALATE/F_BLATE; shared memory): byte offset of the result-
                         ring entry | 1 (the kernel adds thread*8), or of the constant-pool entry
    word 3   byte offset slot*256 of the result in the scratchbook (when F_STORE)

``op == OP_MADD`` is the hot path: ``acc = [acc] +- A*B`` (F_FIRST drops the old accumulator,
F_NEGA negates the product); it covers sums of products of pairs, i.e. the dot products of the
LDL factorisation, the triangular solves and the mat-vec products.  Every other cell kind goes
through one out-of-line switch.

An instruction of the reference ISA becomes 1..k cells; the LAST one
carries F_END: the accumulator is the instruction's value, it is pushed into the per-lane
RESULT RING (shared memory, the last RING results of the group) and, when F_STORE, written to
the scratchbook.  The evaluation order inside an instruction is the C expression order of the
reference emitter and the kernels round every multiply and add separately, so results are
bit-identical to the CPU oracle wherever no libm function is involved.

The kernels fetch scratchbook operands one BUNDLE (``BUNDLE`` cells) ahead of execution to
keep ~2*BUNDLE loads in flight per warp.  That is only legal for operands that are not
produced inside the look-ahead window: the lowering below redirects every operand whose
producer is one of the previous ``RING`` results of the same group to the result ring, which
covers the window (every instruction has at least one cell, and RING > 2*BUNDLE).  Values whose
every use is served by the ring are never stored to the scratchbook at all.  Constants live in
shared memory next to the ring.

Image = 64 int64 header + 16 double header + sections (see ``serialize``).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from . import itable as it
from .csparse import CompiledProblem
from .optimize import GeneratedSolver

MAGIC = 0x3030324243540006  # "\x03\0TCB200" little-endian tag + format version
CONST_BIT = 1 << 30         # io[] words only (get/copy sources): constant-pool index
NO_SLOT = 0xFFFFFFFF

BUNDLE = 4                  # cells fetched/decoded together -- MUST match TC_BUNDLE in csrc/tc_image.h
RING = 32                   # results kept in the shared-memory ring -- MUST match TC_RING
BLOCK = 128                 # threads per CTA -- MUST match TC_BLOCK
MAX_CONSTS = 1024           # constant pool entries held in shared memory -- MUST match TC_MAXCONST
TILE_BYTES = 256            # one scratchbook slot of one warp: 32 lanes x 8 bytes
RING_BYTES = RING * BLOCK * 8

F_ALATE = 1 << 7            # operand A is read from shared memory at execution time
F_BLATE = 1 << 8
F_STORE = 1 << 9            # write the result to the scratchbook at byte offset word3
F_END = 1 << 10             # last cell of an instruction: push acc to the ring
F_NEG = 1 << 11             # result = -acc
F_NEGA = 1 << 12            # operand A (OP_MADD: the product) enters with a minus sign
F_NEGB = 1 << 13            # operand B enters with a minus sign (sums)
F_INIT = 1 << 14            # min0/max0: acc=0 first; clp: acc=DBL_MAX first
F_FIRST = 1 << 15           # OP_MADD: acc = +-A*B (no accumulation)

# cell opcodes -- MUST match enum tc_cellop in csrc/tc_image.h
(OP_MADD, OP_SUM1F, OP_SUM2F, OP_SUM1, OP_SUM2, OP_NOP, OP_RESERVED, OP_PSET, OP_PMUL1, OP_PMUL2,
 OP_P1SET, OP_P1ADD, OP_P2SET, OP_P2ADD, OP_RSUB, OP_RSUBDIV, OP_NEGDIV, OP_DIV, OP_SQ1F, OP_SQ2F,
 OP_SQ1, OP_SQ2, OP_ABS1F, OP_ABS2F, OP_ABS1, OP_ABS2, OP_MINF1, OP_MINF2, OP_MIN1, OP_MIN2,
 OP_MAXF1, OP_MAXF2, OP_MAX1, OP_MAX2, OP_AMAXF1, OP_AMAXF2, OP_AMAX1, OP_AMAX2, OP_CLP, OP_PACC) = range(40)
OP_UNARY = 64               # + reference opcode of the unary function (I_inv ... I_heaviside)

# driver function ids -- MUST match enum tc_fn in csrc/tc_image.h
FN = {name: k for k, name in enumerate([
    "initPrimal__", "scaleIneq__", "scaleCost__", "getScale4Cost__", "setMu__", "initDualIneq__",
    "initDualEqX__", "setAddEye2HessianU__", "setAddEye2HessianEq__", "getNorminf_Grad__",
    "getNorminf_G__", "getGapMinFMinLambda__", "getCurvature__", "getDirectionError__",
    "getHessInertia__", "getMaxAlphas_s__", "setAlphaPrimal__", "getMinF_s__", "setAlphaDualEq__",
    "setAlphaDualIneq__", "updatePrimalDual__", "getJ__", "getOutputs",
    "getMaxAlphas_a__", "getMinF_a__", "getRho__",
    "setAddEye2HessianD__", "getHessDinertia__", "getHessUinertia__",
    "saveWW__"])}               # allowSave (ipmPD_CSsolver.c:51): copy of the KKT values into their save variables
FN_USER0 = len(FN)
# the equilibrium driver (lib/ipmPDeq_CSsolver.c:28-52) calls three of the slots by other names
EQ_NAMES = {"setAddEye2HessianU__": "setAddEye2Hessian1__", "setAddEye2HessianEq__": "setAddEye2Hessian2__",
            "initDualEqX__": "initDualEq__"}
DRIVERS = {"ipmPD_CSsolver": 0, "ipmPDeq_CSsolver": 1, "ipmPDminmax_CSsolver": 2}


def driver_of(pr: CompiledProblem) -> str:
    return str((pr.defines or {}).get("driver", "ipmPD_CSsolver"))


def fn_items(pr: CompiledProblem):
    """(name of the function in THIS problem, driver function id) for every slot of the FN table"""
    eq = driver_of(pr) == "ipmPDeq_CSsolver"
    return [((EQ_NAMES.get(name, name) if eq else name), fid) for name, fid in FN.items()]
KIND_NONE, KIND_SET, KIND_GET, KIND_COPY = 0, 1, 2, 3


GW = 4                      # 64-bit words of a dependency-group mask -- MUST match TC_GW
MAX_GROUPS = 64 * GW


def _mask(groups) -> List[int]:
    words = [0] * GW
    for g in groups:
        if g >= MAX_GROUPS:
            raise ValueError(f"more than {MAX_GROUPS} dependency groups")
        words[g >> 6] |= 1 << (g & 63)
    return words


def persistent_state(pr: CompiledProblem):
    """What must move with an instance when the phased solve compacts the unfinished instances
    (csrc/tc_driver.cuh, tc_abi.cu): the slots of every variable (``I_set``: destinations of set functions
    and copies) and of the values that STATIC groups -- groups no solver-side set or copy invalidates,
    i.e. parameter-only subexpressions -- export to other groups, gets or copies.  Returns
    (sorted slot list, static group ids)."""
    tab = pr.tab
    dynamic = set()
    for s in pr.sets.values():
        if s.kind == "solver":
            dynamic.update(s.children)
    for c in pr.copies.values():
        dynamic.update(c.children)
    static = [g for g in range(len(pr.groups)) if g not in dynamic]
    ext = set()
    for g in pr.gets.values():
        ext.update(x for s in g.sources for x in s if x is not None)
    for c in pr.copies.values():
        ext.update(c.src)
    for i, gi in pr.group_of.items():
        for o in tab.operands[i]:
            if pr.group_of.get(o, -1) != gi:
                ext.add(o)
    slots = {int(pr.slot[i]) for i in range(len(tab)) if tab.types[i] == it.I_set and pr.slot[i] >= 0}
    for g in static:
        slots.update(int(pr.slot[i]) for i in pr.groups[g] if i in ext and pr.slot[i] >= 0)
    return sorted(slots), static


SELF = ("self", 0)           # operand token: the ring entry of the instruction being lowered


def instruction_cells(t: int, p, opw: List[int], compiled: bool = False) -> List[List[int]]:
    """Cells [op, flags, A, B] of one instruction; ``opw`` are operand tokens in the reference's
    operand order (``None`` marks an unused operand).  The arithmetic order is that of the emitted C statement
    (instructionsTableUTHash.c:950-1539)."""
    N = None
    c: List[List[int]] = []

    def two(xs, f1, f2, n1, n2, flags_first=0):
        """generic left-to-right fold over a list, two operands per cell"""
        for k in range(0, len(xs), 2):
            pair = xs[k:k + 2]
            first = k == 0
            if len(pair) == 2:
                c.append([f2 if first else n2, flags_first if first else 0, pair[0], pair[1]])
            else:
                c.append([f1 if first else n1, flags_first if first else 0, pair[0], N])

    def dot(xs):
        for k in range(0, len(xs), 2):
            c.append([OP_MADD, F_FIRST if k == 0 else 0, xs[k], xs[k + 1]])

    if t in (it.I_load, it.I_2):
        c.append([OP_SUM1F, 0, opw[0], N])
    elif t == it.I_sum:                                          # +a-b+c...           :964-968
        for k in range(0, len(opw), 2):
            pair = opw[k:k + 2]
            fl = (F_NEGA if p[k] < 0 else 0) | (F_NEGB if len(pair) == 2 and p[k + 1] < 0 else 0)
            if len(pair) == 2:
                c.append([OP_SUM2F if k == 0 else OP_SUM2, fl, pair[0], pair[1]])
            else:
                c.append([OP_SUM1F if k == 0 else OP_SUM1, fl, pair[0], N])
    elif t == it.I_sumprod:                                      # a*b+c*d+...         :1007-1029
        m = int(p[0])
        per = m * -(-it.MAX_TERMS_PERLINE // m)                  # operands per emitted line (it.sumprod_lines)
        if len(opw) > per:
            # `m[d]=line1; m[d]+=line2; ...` (:1019-1026): every line is an accumulation of its own that ends in
            # the instruction's ring entry; the following line adds that entry (a late operand) to its own sum
            for k in range(0, len(opw), per):
                c.extend(instruction_cells(t, p, opw[k:k + per]))
                if k:
                    c[-1][1] &= ~F_END
                    c.append([OP_SUM1, F_END, SELF, N])
            return c
        if m == 1:
            two(opw, OP_SUM1F, OP_SUM2F, OP_SUM1, OP_SUM2)
        elif m == 2:
            dot(opw)
        else:
            for k in range(0, len(opw), m):
                f = opw[k:k + m]
                first = k == 0
                c.append([OP_PSET, 0, f[0], f[1]])
                rest = f[2:]
                while len(rest) > 2:
                    c.append([OP_PMUL2, 0, rest[0], rest[1]])
                    rest = rest[2:]
                if len(rest) == 2:
                    c.append([OP_P2SET if first else OP_P2ADD, 0, rest[0], rest[1]])
                else:
                    c.append([OP_P1SET if first else OP_P1ADD, 0, rest[0], N])
    elif t == it.I_plus_minus_dot_div:                           # (a0-(dot))/d        :1081-1098
        dot(opw[2:])
        c.append([OP_RSUBDIV, 0, opw[0], opw[1]])
    elif t == it.I_minus_dot_div:                                # -(dot)/d            :1106-1121
        dot(opw[1:])
        c.append([OP_NEGDIV, 0, opw[0], N])
    elif t == it.I_plus_minus_dot:                               # a0-(dot)            :1128-1144
        dot(opw[1:])
        c.append([OP_RSUB, 0, opw[0], N])
    elif t == it.I_minus_dot:                                    # -(dot)              :1151-1164
        dot(opw)
        c[-1][1] |= F_NEG
    elif t == it.I_div:
        c.append([OP_DIV, 0, opw[0], opw[1]])
    elif t == it.I_plus_sqr:                                     # a*a+b*b+...         :1201-1210
        two(opw, OP_SQ1F, OP_SQ2F, OP_SQ1, OP_SQ2)
    elif t == it.I_plus_abs:                                     # |a|+|b|+...         :1217-1225
        two(opw, OP_ABS1F, OP_ABS2F, OP_ABS1, OP_ABS2)
    elif t == it.I_min:                                          # compare chains      :1228-1313
        two(opw, OP_MINF1, OP_MINF2, OP_MIN1, OP_MIN2)
    elif t == it.I_max:
        two(opw, OP_MAXF1, OP_MAXF2, OP_MAX1, OP_MAX2)
    elif t == it.I_min0:                                         # acc=0, then the chain
        two(opw, OP_MIN1, OP_MIN2, OP_MIN1, OP_MIN2, flags_first=F_INIT)
    elif t == it.I_max0:
        two(opw, OP_MAX1, OP_MAX2, OP_MAX1, OP_MAX2, flags_first=F_INIT)
    elif t == it.I_max_abs:
        two(opw, OP_AMAXF1, OP_AMAXF2, OP_AMAX1, OP_AMAX2)
    elif t == it.I_clp:                                          # :1315-1331
        for k in range(0, len(opw), 2):
            c.append([OP_CLP, F_INIT if k == 0 else 0, opw[k], opw[k + 1]])
    elif t in it.UNARY_TYPES:
        c.append([OP_UNARY + t, 0, opw[0], N])
    elif t == it.I_componentwise:
        # user C text: only the compiled module can run it.  An image made for the compiled back end carries a
        # placeholder cell (the interpreter evaluates unknown unary opcodes to NaN: loud, never silently wrong)
        if not compiled:
            raise NotImplementedError("I_componentwise is user C text: it needs the compiled back end "
                                      "(backend='compiled'), the tape interpreter cannot run it")
        c.append([OP_UNARY + t, 0, opw[0], N])
    else:
        raise NotImplementedError(it.OPNAMES.get(t, str(t)))
    c[-1][1] |= F_END
    return c


def lower(gen: GeneratedSolver, ring: int = RING, bundle: int = BUNDLE, compiled: bool = False) -> Dict[str, np.ndarray]:
    """``compiled``: the image is for a handle that runs the problem's compiled module -- instructions only the
    module can evaluate (``I_componentwise``) are accepted and lowered to placeholder cells"""
    pr: CompiledProblem = gen.problem
    tab, slot = pr.tab, pr.slot
    if pr.n_slots * TILE_BYTES >= 1 << 32:
        raise ValueError("scratchbook too large for 32-bit tile offsets")
    if ring <= 2 * bundle:
        raise ValueError("the result ring must cover the operand look-ahead window (RING > 2*BUNDLE)")
    consts: List[float] = []
    cindex: Dict[float, int] = {}

    def cidx(v: float) -> int:
        k = cindex.get(v)
        if k is None:
            k = cindex[v] = len(consts)
            consts.append(v)
        return k

    def const_of(i: int):
        t = tab.types[i]
        if t == it.I_load:
            return tab.params[i][0]
        if t == it.I_2:
            return 2.0
        return None

    def io_word(i: int) -> int:                 # get/copy sources: slot or constant
        v = const_of(i)
        return (CONST_BIT | cidx(v)) if v is not None else int(slot[i])

    # values that must reach the scratchbook: read by another group, a get or a copy
    external = set()
    for g in pr.gets.values():
        external.update(x for s in g.sources for x in s if x is not None)
    for c in pr.copies.values():
        external.update(c.src)
    for i, gi in pr.group_of.items():
        for o in tab.operands[i]:
            if pr.group_of.get(o, -1) != gi:
                external.add(o)

    cells: List[List[int]] = []                 # [w0, w1, w2, w3]
    self_ring = [0]
    group_start = [0]
    stats = dict(slot_reads=0, ring_reads=0, const_reads=0, stores=0, ring_only=0)

    def enc(tok, late_flag: int) -> Tuple[int, int]:
        """operand token -> (word, flag)"""
        if tok is None:                         # unused operand: a harmless shared-memory read
            return RING_BYTES, late_flag
        kind, v = tok
        if kind == "slot":
            return v * TILE_BYTES, 0
        if kind == "self":
            kind, v = "ring", self_ring[0]
        if kind == "ring":
            return (v * BLOCK * 8) | 1, late_flag
        return RING_BYTES + 8 * v, late_flag    # constant pool, stored right after the ring

    for members in pr.groups:
        order = {i: n for n, i in enumerate(members)}
        # furthest in-group consumer of every result, in results
        far = {}
        for n, i in enumerate(members):
            for o in tab.operands[i]:
                k = order.get(o)
                if k is not None:
                    far[o] = max(far.get(o, 0), n - k)
        for n, i in enumerate(members):
            t, p, ops = tab.types[i], tab.params[i], tab.operands[i]
            if t in (it.I_load, it.I_2):
                opw = [("const", cidx(const_of(i)))]
                stats["const_reads"] += 1
            else:
                opw = []
                for o in ops:
                    v = const_of(o)
                    k = order.get(o)
                    if v is not None:
                        opw.append(("const", cidx(v)))
                        stats["const_reads"] += 1
                    elif k is not None and n - k < ring:
                        opw.append(("ring", k % ring))
                        stats["ring_reads"] += 1
                    else:
                        opw.append(("slot", int(slot[o])))
                        stats["slot_reads"] += 1
            cs = instruction_cells(t, p, opw, compiled)
            store = i in external or far.get(i, 0) >= ring
            stats["stores" if store else "ring_only"] += 1
            self_ring[0] = n % ring
            for q, (op, fl, a, b_) in enumerate(cs):
                wa, fa = enc(a, F_ALATE)
                wb, fb = enc(b_, F_BLATE)
                w0 = op | fl | fa | fb
                w3 = 0
                if fl & F_END:                  # (a multi-line sumprod ends every line in its ring entry)
                    w0 |= (n % ring) << 16
                if q == len(cs) - 1:
                    w0 |= F_STORE if store else 0
                    w3 = int(slot[i]) * TILE_BYTES if store else 0
                cells.append([w0, wa, wb, w3])
        while len(cells) % bundle:
            cells.append([OP_NOP | F_ALATE | F_BLATE, RING_BYTES, RING_BYTES, 0])
        group_start.append(len(cells))
    # more than MAX_CONSTS distinct constants: fine for the compiled back end (constants are immediates there);
    # the interpreter keeps the pool in shared memory and refuses such an image at solve time (tc_abi.cu)

    # ---- function table ------------------------------------------------------------------------
    user_sets = [s for s in pr.sets.values() if s.kind in ("parameter", "variable")]
    n_funcs = FN_USER0 + len(user_sets)
    funcs = np.zeros((n_funcs, 2 * GW + 2), dtype=np.uint64)   # parents[GW], children[GW], io_off|io_cnt<<32, kind
    io: List[int] = []

    def put(fid: int, kind: int, parents, children, words: List[int]) -> None:
        funcs[fid] = tuple(_mask(parents)) + tuple(_mask(children)) + (len(io) | (len(words) << 32), kind)
        io.extend(words)

    def src_word(x) -> int:
        return NO_SLOT if x is None else io_word(x)

    for name, fid in fn_items(pr):
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

    persist, static = persistent_state(pr)
    return dict(persist=np.asarray(persist if persist else [0], dtype=np.uint32), n_persist=len(persist),
                static_mask=_mask(static),
                tape=np.asarray(cells, dtype=np.uint32).reshape(-1, 4),
                group_start=np.asarray(group_start, dtype=np.uint32),
                consts=np.asarray(consts if consts else [0.0], dtype=np.float64), funcs=funcs,
                io=np.asarray(io if io else [0], dtype=np.uint32),
                set_fids=np.asarray(set_fids, dtype=np.uint32), out_sizes=np.asarray(out_sizes, dtype=np.uint32),
                stats=stats)


def serialize(gen: GeneratedSolver, compiled: bool = False) -> bytes:
    L = lower(gen, compiled=compiled)
    d = gen.defines
    ih = np.zeros(64, dtype=np.int64)
    ih[0] = MAGIC
    ih[1:9] = (gen.problem.n_slots, len(gen.problem.groups), len(L["funcs"]), len(L["set_fids"]),
               len(L["out_sizes"]), len(L["tape"]), len(L["consts"]), len(L["io"]))
    ih[9:13] = (RING, BUNDLE, L["n_persist"], GW)
    ih[32:32 + GW] = np.asarray(L["static_mask"], dtype=np.uint64).view(np.int64)
    ih[16:19] = (d["nU"], d["nG"], d["nF"])
    ih[19:27] = (d["scaleCost"], d["scaleInequalities"], d["coupledAlphas"], d["setAddEye2Hessian"],
                 d["adjustAddEye2Hessian"], d["useInertia"], d["skipAffine"], d["delta"])
    ih[27] = DRIVERS[str(d.get("driver", "ipmPD_CSsolver"))]
    if ih[27] in (1, 2):
        ih[16] = d["nZ"]          # the device driver only tests nG > 0 / nF > 0 and divides by nF
    if ih[27] == 2:
        ih[28:30] = (d["mpUdesired"], d["mnDdesired"])
    ih[30] = int(d.get("allowSave", 0))
    dh = np.zeros(16, dtype=np.float64)
    dh[0:8] = (d["gradTolerance"], d["addEye2HessianUtolerance"], d["equalTolerance"], d["desiredDualityGap"],
               d["alphaMin"], d["alphaMax"], d["muFactorAggressive"], d["muFactorConservative"])
    dh[8] = d.get("addEye2HessianDtolerance", 0.0)

    def pad16(b: bytes) -> bytes:
        return b + b"\0" * (-len(b) % 16)

    # names of the set functions (the <name> of the reference's `<class>_set_<name>` symbols, parameters first, then
    # optimisation variables: writeCfunctionpergroup.m:524) and of the outputs, NUL-terminated, in index order --
    # what tc_find_set / tc_find_output look up
    names = [n for n, _ in gen.parameters] + [n for n, _ in gen.variables] + [n for n, _ in gen.outputs]
    blob = b"".join(n.encode() + b"\0" for n in names)
    if len(names) != len(L["set_fids"]) + len(L["out_sizes"]) or any(not n or "\0" in n for n in names):
        raise ValueError("set/output names do not match the set functions and outputs of the problem")
    ih[13] = len(blob)
    parts = [ih.tobytes(), dh.tobytes()]
    for key in ("group_start", "tape", "consts", "funcs", "io", "set_fids", "out_sizes", "persist"):
        parts.append(pad16(np.ascontiguousarray(L[key]).tobytes()))
    parts.append(pad16(blob))
    return b"".join(parts)
