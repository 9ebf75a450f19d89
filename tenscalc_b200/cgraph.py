"""Computation-graph files ``.cg / .cgc / .cgio / .cgs``: the reference's own wire format for the
scalar instruction graph of a csparse object (``lib/@csparse/saveScalarized.m:108-273``, opcodes
``lib/CGregistration.m:65-133``, format ``doc/computationgraphs.tex:84-205``).

``load_scalarized`` ingests what a MATLAB user writes with ``saveScalarized(code,'name')`` for a
solver built by ``cmex2optimizeCS`` and returns the same :class:`GeneratedSolver` the Python front
end produces -- instruction table, set/get/copy declarations with the reference's function names,
dependency groups -- ready for ``gen(batch)`` on the GPU.  No re-derived front end is involved: the
instruction list is the one MATLAB built.  ``save_scalarized`` writes the same four files from a
:class:`GeneratedSolver` (MATLAB is not available here, so the round trip through this writer is
what the tests pin; every record layout below cites the line of saveScalarized.m that writes it).

File layout as the reference WRITES it (where the code and the .tex differ, the code wins):

* every file starts with ONE byte: the magic 16/32/64 = width of every integer that follows,
  little endian (``fwrite(fg,magic,'uchar')`` :117-120; saveScalarized uses 32, :15);
* ``.cg``  function nodes.  ``[0, const#, outvar]`` for a constant node, otherwise
  ``[type, nIn, nOut, in..., out...]``, 0-based variables (:166,172,181,146).  Scalar instruction k
  of the table is variable k; ``I_set`` writes nothing (:174-176); ``I_sum``/``I_sumprod`` get one
  extra leading input, a constant variable holding their parameter vector (:168-172);
* ``.cgc`` constants ``[type, nbytes, data]``: 10001 double tensor ``[nd, size.., values]``,
  20001 int64 vector, 20002 int64 tensor ``[nd, size.., values]``, 30001 string (:229-258);
* ``.cgio`` gateway functions ``[2=get, n, vars]``, ``[1=set, n, vars]``,
  ``[3=copy, n, src.., dst..]`` (:199,219,244).  NOTE the reference writes the get/set variables
  1-BASED there (``newvars`` instead of ``newvars-1``) although the matching ``.cg`` nodes and the
  copies are 0-based; the reader detects which convention a file uses;
* ``.cgs`` symbols ``[kind, index, len(name), len(description)]`` + chars; kind 4 = gateway
  function names (:200-201).

The reference registers only the opcodes ``Ssum..SDDatan`` (20001..20032); instruction types
without a registration (``I_2_inv_cube``, ``I_plus_abs``, ``I_round`` ...) make its writer stop
with "missing type" (:183).  ``save_scalarized(..., extended=True)`` stores those under private
codes 29000+opcode so that every problem of this repo can be saved; the reader accepts both.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import itable as it
from .csparse import CSparse
from .itable import InstructionTable

# lib/CGregistration.m:65-133
CG_CONST_FUNCTION = 0
CG_S2FT, CG_FT2S = 30001, 30002
CG_IO_SET, CG_IO_GET, CG_IO_COPY = 1, 2, 3
CG_SYM_CONST, CG_SYM_FUNCTION, CG_SYM_VARIABLE, CG_SYM_IO = 1, 2, 3, 4
CGC_DOUBLE, CGC_INT_VECTOR, CGC_INT_TENSOR, CGC_STRING = 10001, 20001, 20002, 30001
SCALAR_CODES = {
    it.I_sum: 20001, it.I_sumprod: 20002, it.I_inv: 20003, it.I_minus_inv_sqr: 20004, it.I_div: 20005,
    it.I_minus_dot: 20006, it.I_minus_dot_div: 20007, it.I_plus_minus_dot: 20008, it.I_plus_sqr: 20009,
    it.I_plus_minus_dot_div: 20010, it.I_min: 20011, it.I_min0: 20012, it.I_max: 20013, it.I_max0: 20014,
    it.I_max_abs: 20015, it.I_clp: 20016, it.I_exp: 20017, it.I_log: 20018, it.I_cos: 20019,
    it.I_minus_cos: 20020, it.I_sin: 20021, it.I_minus_sin: 20022, it.I_abs: 20023, it.I_sqrt: 20024,
    it.I_Dsqrt: 20025, it.I_DDsqrt: 20026, it.I_sqr: 20027, it.I_2times: 20028, it.I_2: 20029,
    it.I_atan: 20030, it.I_Datan: 20031, it.I_DDatan: 20032,
}
EXT_BASE = 29000                      # private: 29000 + reference opcode, for unregistered instruction types
CODE_TO_OP = {v: k for k, v in SCALAR_CODES.items()}


class CGFormatError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self, base: str, width: int):
        self.fmt = {16: "<h", 32: "<i", 64: "<q"}[width]
        self.f = {ext: open(f"{base}.{ext}", "wb") for ext in ("cg", "cgc", "cgio", "cgs")}
        for f in self.f.values():
            f.write(bytes([width]))
        self.constants: List[Tuple[object, str, str]] = []
        self.n_functions = 0

    def ints(self, ext: str, values: Sequence[int]) -> None:
        self.f[ext].write(b"".join(struct.pack(self.fmt, int(v)) for v in values))

    def symbol(self, kind: int, index: int, name: str, description: str) -> None:
        self.ints("cgs", [kind, index, len(name), len(description)])
        self.f["cgs"].write((name + description).encode("latin-1"))

    def add_constant(self, value, name: str = "", description: str = "") -> int:
        self.constants.append((value, name, description))
        return len(self.constants) - 1

    def close(self) -> None:
        for f in self.f.values():
            f.close()


def save_scalarized(gen, filename: str, extended: bool = True, int_width: int = 32) -> Dict[str, int]:
    """Write ``filename.cg/.cgc/.cgio/.cgs`` for a generated solver, record for record as
    saveScalarized.m does (including its 1-based get/set variables in ``.cgio``)."""
    if os.path.splitext(filename)[1]:
        raise ValueError(f'filename "{filename}" must not include extension')      # :113-115
    pr = gen.problem
    tab: InstructionTable = pr.tab
    w = _Writer(filename, int_width)
    n_var = len(tab)                                     # instruction k is variable k; extras follow

    def constant_variable(value, description: str) -> int:   # addConstantVariable :27-48
        nonlocal n_var
        var = n_var
        n_var += 1
        const = w.add_constant(value, "", description)
        w.ints("cg", [CG_CONST_FUNCTION, const, var])
        w.symbol(CG_SYM_FUNCTION, w.n_functions, "", description)
        w.symbol(CG_SYM_VARIABLE, var, "", description)
        w.n_functions += 1
        return var

    for i in range(len(tab)):
        t, p, ops = tab.types[i], tab.params[i], tab.operands[i]
        if t == it.I_set:                                # :174-176
            continue
        if t == it.I_load:                               # :178-181
            const = w.add_constant(np.array([[float(p[0])]]), "", "load")
            w.ints("cg", [CG_CONST_FUNCTION, const, i])
        elif t in (it.I_sum, it.I_sumprod):              # :166-172
            var = constant_variable(np.asarray(p, dtype=np.int64).reshape(-1, 1), "sum")
            w.ints("cg", [SCALAR_CODES[t], 1 + len(ops), 1, var, *ops, i])
        elif t == it.I_componentwise:                    # private extension: the C text travels as a parameter vector
            if not extended:                             # (saveScalarized has no code for it, :138-195)
                raise CGFormatError(f"csparse/saveScalarized: missing type {t} ({it.OPNAMES.get(t)})")
            var = constant_variable(np.asarray(p, dtype=np.int64).reshape(-1, 1), "componentwise")
            w.ints("cg", [EXT_BASE + t, 1 + len(ops), 1, var, *ops, i])
        else:
            code = SCALAR_CODES.get(t)
            if code is None:
                if not extended:
                    raise CGFormatError(f"csparse/saveScalarized: missing type {t} ({it.OPNAMES.get(t)})")   # :183
                code = EXT_BASE + t
            w.ints("cg", [code, len(ops), 1, *ops, i])   # :146
        w.n_functions += 1

    n_io = 0

    def io_name(name: str) -> str:
        if name.startswith(("setP_", "setV_")):
            return f"{gen.classname}_set_{name[5:]}"
        if name == "getOutputs":
            return f"{gen.classname}_getOutputs"
        return name

    for g in pr.gets.values():                           # :188-203
        newvars = list(range(n_var, n_var + len(g.sources)))   # the full tensors first, then their subscripts
        n_var += len(g.sources)
        for newvar, src, shape in zip(newvars, g.sources, g.shapes):
            dims = tuple(shape)
            nz = [(k, x) for k, x in enumerate(src) if x is not None]
            if dims:
                subs = np.array(np.unravel_index([k for k, _ in nz], dims, order="F"), dtype=np.int64).reshape(len(dims), -1) + 1
            else:
                subs = np.zeros((0, len(nz)), dtype=np.int64)
            var = constant_variable(subs, "get")
            w.ints("cg", [CG_S2FT, 1 + len(nz), 1, var, *[x for _, x in nz], newvar])
            w.n_functions += 1
        w.ints("cgio", [CG_IO_GET, len(newvars), *[v + 1 for v in newvars]])        # 1-based, as the reference writes
        w.symbol(CG_SYM_IO, n_io, io_name(g.name), "")
        n_io += 1
    for s in pr.sets.values():                           # :204-223 (one destination tensor per set here)
        newvar = n_var
        n_var += 1
        w.ints("cg", [CG_FT2S, 1, len(s.dest), newvar, *s.dest])
        w.n_functions += 1
        w.ints("cgio", [CG_IO_SET, 1, newvar + 1])
        w.symbol(CG_SYM_IO, n_io, io_name(s.name), "")
        n_io += 1
    for c in pr.copies.values():                         # :224-250
        w.ints("cgio", [CG_IO_COPY, len(c.src), *c.src, *c.dest])
        w.symbol(CG_SYM_IO, n_io, io_name(c.name), "")
        n_io += 1

    fc = w.f["cgc"]
    for k, (value, name, description) in enumerate(w.constants):      # :252-283
        if isinstance(value, str):
            w.ints("cgc", [CGC_STRING, len(value)])
            fc.write(value.encode("latin-1"))
        elif value.dtype == np.int64:
            if value.ndim == 2 and value.shape[1] == 1:
                w.ints("cgc", [CGC_INT_VECTOR, 8 * value.size])
                fc.write(np.asarray(value, dtype="<i8").tobytes(order="F"))
            else:
                sz = value.shape
                w.ints("cgc", [CGC_INT_TENSOR, 8 * (1 + len(sz) + value.size)])
                fc.write(np.array([len(sz), *sz], dtype="<i8").tobytes())
                fc.write(np.asarray(value, dtype="<i8").tobytes(order="F"))
        else:
            sz = value.shape
            w.ints("cgc", [CGC_DOUBLE, 8 * (1 + len(sz) + value.size)])
            fc.write(np.array([len(sz), *sz], dtype="<i8").tobytes())
            fc.write(np.asarray(value, dtype="<f8").tobytes(order="F"))
        if name or description:
            w.symbol(CG_SYM_CONST, k, name, description)
    w.close()
    return dict(functions=w.n_functions, ios=n_io, constants=len(w.constants), variables=n_var)


# ---------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            data = f.read()
        if not data or data[0] not in (16, 32, 64):
            raise CGFormatError(f"{path}: bad magic")
        self.width = data[0] // 8
        self.dtype = {2: "<i2", 4: "<i4", 8: "<i8"}[self.width]
        self.data = data
        self.pos = 1
        self.path = path

    def eof(self) -> bool:
        return self.pos >= len(self.data)

    def ints(self, n: int) -> List[int]:
        end = self.pos + n * self.width
        if n < 0 or end > len(self.data):
            raise CGFormatError(f"{self.path}: truncated record at byte {self.pos}")
        out = np.frombuffer(self.data, dtype=self.dtype, count=n, offset=self.pos)
        self.pos = end
        return [int(x) for x in out]

    def raw(self, n: int) -> bytes:
        end = self.pos + n
        if n < 0 or end > len(self.data):
            raise CGFormatError(f"{self.path}: truncated data at byte {self.pos}")
        out = self.data[self.pos:end]
        self.pos = end
        return out


def read_files(filename: str):
    """parse the four files into plain records: (constants, nodes, ios, io_names)"""
    rc = _Reader(filename + ".cgc")
    constants: List[object] = []
    while not rc.eof():
        ctype, nbytes = rc.ints(2)
        blob = rc.raw(nbytes)
        if ctype == CGC_STRING:
            constants.append(blob.decode("latin-1"))
        elif ctype == CGC_INT_VECTOR:
            constants.append(np.frombuffer(blob, dtype="<i8").reshape(-1, 1).copy())
        elif ctype in (CGC_INT_TENSOR, CGC_DOUBLE):
            head = np.frombuffer(blob, dtype="<i8", count=1)
            nd = int(head[0])
            sz = [int(x) for x in np.frombuffer(blob, dtype="<i8", count=nd, offset=8)]
            dt = "<i8" if ctype == CGC_INT_TENSOR else "<f8"
            vals = np.frombuffer(blob, dtype=dt, offset=8 * (1 + nd))
            if vals.size != int(np.prod(sz)):
                raise CGFormatError(f"{filename}.cgc: constant {len(constants)} has {vals.size} values for size {sz}")
            constants.append(vals.reshape(sz, order="F").copy())
        else:
            raise CGFormatError(f"{filename}.cgc: unregistered constant type {ctype}")
    rg = _Reader(filename + ".cg")
    nodes: List[Tuple[int, List[int], List[int]]] = []     # (type, inputs, outputs); constant: (0, [const#], [out])
    while not rg.eof():
        (t,) = rg.ints(1)
        if t == CG_CONST_FUNCTION:
            c, out = rg.ints(2)
            if not 0 <= c < len(constants):
                raise CGFormatError(f"{filename}.cg: constant record {c} does not exist")
            nodes.append((t, [c], [out]))
        else:
            n_in, n_out = rg.ints(2)
            nodes.append((t, rg.ints(n_in), rg.ints(n_out)))
    rio = _Reader(filename + ".cgio")
    ios: List[Tuple[int, List[int], List[int]]] = []
    while not rio.eof():
        kind, n = rio.ints(2)
        if kind == CG_IO_COPY:
            ios.append((kind, rio.ints(n), rio.ints(n)))
        elif kind in (CG_IO_SET, CG_IO_GET):
            ios.append((kind, rio.ints(n), []))
        else:
            raise CGFormatError(f"{filename}.cgio: unknown gateway type {kind}")
    io_names: Dict[int, str] = {}
    if os.path.exists(filename + ".cgs"):
        rs = _Reader(filename + ".cgs")
        while not rs.eof():
            kind, index, ln, ld = rs.ints(4)
            text = rs.raw(ln + ld).decode("latin-1")
            if kind == CG_SYM_IO:
                io_names[index] = text[:ln]
    return constants, nodes, ios, io_names


def load_scalarized(filename: str, classname: Optional[str] = None, output_shapes: Optional[Dict[str, list]] = None,
                    **solver_options):
    """Read ``filename.cg/.cgc/.cgio/.cgs`` (written by the reference's ``saveScalarized`` for the csparse
    object of a ``cmex2optimizeCS`` solver, or by :func:`save_scalarized`) into a
    :class:`~tenscalc_b200.optimize.GeneratedSolver`.  ``solver_options`` are the name/value options of
    ``cmex2optimizeCS`` that the files do not carry (tolerances, ``muFactorAggressive`` ...: the
    reference bakes them into the C driver as #defines, cmex2optimizeCS.m:303-331); sizes and the
    option switches that change the declared functions (nU, nG, nF, scaleCost, scaleInequalities,
    useInertia) are recovered from the gateway functions themselves.  ``output_shapes`` =
    ``{"getOutputs": [shape, ...]}`` gives the full sizes of get tensors whose trailing rows/columns are
    structurally zero (the files only hold non-zero subscripts)."""
    from .optimize import GeneratedSolver, solver_defines
    constants, nodes, ios, io_names = read_files(filename)

    produced: Dict[int, Tuple[int, List[int], List[int]]] = {}
    for node in nodes:
        for o in node[2]:
            if o in produced and node[0] != CG_FT2S:
                raise CGFormatError(f"variable {o} is the output of two function nodes")
            produced.setdefault(o, node)
    # auxiliary variables (parameter vectors, subscripts, full tensors) are numbered after the table
    aux = [n[2][0] for n in nodes if n[0] == CG_CONST_FUNCTION and not _is_load(constants[n[1][0]])]
    aux += [n[2][0] for n in nodes if n[0] == CG_S2FT] + [n[1][0] for n in nodes if n[0] == CG_FT2S]
    used = [v for n in nodes if n[0] not in (CG_CONST_FUNCTION, CG_S2FT, CG_FT2S) for v in n[1][(1 if n[0] in (20001, 20002, EXT_BASE + it.I_componentwise) else 0):] + n[2]]
    used += [n[2][0] for n in nodes if n[0] == CG_CONST_FUNCTION and _is_load(constants[n[1][0]])]
    used += [v for n in nodes if n[0] == CG_S2FT for v in n[1][1:]] + [v for n in nodes if n[0] == CG_FT2S for v in n[2]]
    used += [v for k, a, b in ios if k == CG_IO_COPY for v in a + b]
    n_table = max(used) + 1 if used else 0
    if aux and min(aux) < n_table:
        raise CGFormatError("auxiliary variables overlap the instruction table")

    tab = InstructionTable()
    for v in range(n_table):
        node = produced.get(v)
        if node is None or node[0] == CG_FT2S:
            tab.append(it.I_set)                         # written by a set or a copy (or unused)
            continue
        t, ins, _ = node
        if t == CG_CONST_FUNCTION:
            tab.append(it.I_load, (float(np.asarray(constants[ins[0]]).reshape(-1)[0]),))
            continue
        op = CODE_TO_OP.get(t, t - EXT_BASE if EXT_BASE < t < EXT_BASE + 100 else None)
        if op is None or op not in it.OPNAMES:
            raise CGFormatError(f"unregistered function type {t} for variable {v}")
        params: Tuple[float, ...] = ()
        if op in (it.I_sum, it.I_sumprod, it.I_componentwise):
            pnode = produced.get(ins[0])
            if pnode is None or pnode[0] != CG_CONST_FUNCTION:
                raise CGFormatError(f"variable {v}: parameter input {ins[0]} is not a constant")
            params = tuple(float(x) for x in np.asarray(constants[pnode[1][0]]).reshape(-1))
            ins = ins[1:]
        if any(not 0 <= o < v for o in ins):
            raise CGFormatError(f"variable {v} reads a variable that is not defined before it")
        tab.append(op, params, ins)

    # get/set variables: 0-based per the format description, 1-based as saveScalarized.m writes them
    s2ft = {n[2][0]: n for n in nodes if n[0] == CG_S2FT}
    ft2s = {n[1][0]: n for n in nodes if n[0] == CG_FT2S}
    def resolves(off: int) -> bool:
        return all(v - off in (s2ft if kind == CG_IO_GET else ft2s)
                   for kind, vs, _ in ios if kind != CG_IO_COPY for v in vs)
    if resolves(1):
        off = 1
    elif resolves(0):
        off = 0
    else:
        raise CGFormatError("a get/set gateway names a variable that no tensor node produces")

    names = [io_names.get(k, f"io{k}") for k in range(len(ios))]
    if classname is None:
        cands = [n[:-len("_getOutputs")] for n in names if n.endswith("_getOutputs")]
        classname = cands[0] if cands else os.path.basename(filename)
    code = CSparse(tab)
    copies = {names[k]: (a, b) for k, (kind, a, b) in enumerate(ios) if kind == CG_IO_COPY}
    primal_src = set(copies.get("initPrimal__", ([], []))[0])
    par_info, var_info, out_info = [], [], []
    pending_sets = []
    for k, (kind, vs, dst) in enumerate(ios):
        name = names[k]
        if kind == CG_IO_GET:
            srcs, shapes = [], []
            want = (output_shapes or {}).get(name) or (output_shapes or {}).get(name.replace(classname + "_", ""))
            for j, v in enumerate(vs):
                node = s2ft[v - off]
                subs = np.asarray(constants[produced[node[1][0]][1][0]], dtype=np.int64)
                inst = node[1][1:]
                subs = subs.reshape(-1, len(inst)) if inst else subs.reshape(subs.shape[0] if subs.ndim else 0, 0)
                nd = subs.shape[0]
                if nd == 0:
                    shape, flat = (), list(inst[:1]) or [None]
                else:
                    # the files hold the subscripts of the structural non-zeros only: without `output_shapes`
                    # the tensor is as large as its largest non-zero subscript
                    shape = tuple(want[j]) if want is not None else (tuple(int(x) for x in subs.max(axis=1)) if inst else (0,) * nd)
                    if len(shape) != nd or (inst and any(int(m) > d for m, d in zip(subs.max(axis=1), shape))):
                        raise CGFormatError(f"{name}: output {j} does not fit the shape {shape}")
                    flat = [None] * int(np.prod(shape))
                    lin = np.ravel_multi_index(tuple(subs - 1), shape, order="F") if inst else []
                    for pos, x in zip(lin, inst):
                        flat[int(pos)] = x
                srcs.append(flat)
                shapes.append(shape)
            if name == f"{classname}_getOutputs":
                out_names = [f"y{j + 1}" for j in range(len(srcs))]
                code.declare_get("getOutputs", srcs, shapes, out_names)
                out_info = list(zip(out_names, shapes))
            else:
                code.declare_get(name, srcs, shapes)
        elif kind == CG_IO_SET:
            dest = [x for v in vs for x in ft2s[v - off][2]]
            pending_sets.append((name, dest))
        else:
            code.declare_copy(name, dst, vs)
    for name, dest in pending_sets:
        prefix = classname + "_set_"
        if name.startswith(prefix):
            short = name[len(prefix):]
            shape = (len(dest),) if len(dest) != 1 else ()
            if dest and all(d in primal_src for d in dest):
                code.declare_set("setV_" + short, dest, shape, kind="variable")
                var_info.append((short, shape))
            else:
                code.declare_set("setP_" + short, dest, shape, kind="parameter")
                par_info.append((short, shape))
        else:
            code.declare_set(name, dest)
    for need in ("initPrimal__", "updatePrimalDual__"):
        if need not in code.copies:
            raise CGFormatError(f'the graph declares no "{need}": not the csparse object of an optimisation solver')
    nU = len(code.copies["initPrimal__"].dest)
    nF = len(code.copies["initDualIneq__"].dest) if "initDualIneq__" in code.copies else 0
    nG = len(code.copies["initDualEqX__"].dest) if "initDualEqX__" in code.copies else 0
    inferred = dict(scaleCost=float("scaleCost__" in code.copies), scaleInequalities="scaleIneq__" in code.copies,
                    useInertia="getHessInertia__" in code.gets, addEye2Hessian="setAddEye2HessianU__" in code.sets)
    for k2, v2 in inferred.items():
        if k2 in solver_options and bool(solver_options[k2]) != bool(v2):
            raise CGFormatError(f"option {k2}={solver_options[k2]} contradicts the functions the graph declares")
    defines = solver_defines(nU, nG, nF, **{**solver_options, **inferred})
    problem = code.compile(defines)
    info = dict(ordering="from file", kkt=nU + nG + nF, n_instructions=sum(len(g) for g in problem.groups),
                n_groups=len(problem.groups), n_slots=problem.n_slots, source=filename + ".cg", io_variable_base=off)
    return GeneratedSolver(classname, problem, par_info, var_info, out_info, False, defines, info)


def _is_load(const) -> bool:
    """the constant behind an I_load instruction: a double (saveScalarized.m:179); parameter vectors and
    subscripts are int64"""
    return isinstance(const, np.ndarray) and const.dtype == np.float64
