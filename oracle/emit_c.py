"""CPU ORACLE (test infrastructure only -- never imported by the product path).

Writes one self-contained C translation unit per problem, the counterpart of the file
the reference generates with ``lib/@csparse/writeCfunctionpergroup.m``:

* ``static double scratchbook[S]`` + ``static int groupStatus[G]`` (:148,264-270),
* one straight-line function per dependency group (split every ``MAX_PER_FUNCTION``
  statements, :410-443),
* ``set`` functions (copy + invalidate children, :513-577), ``get`` functions
  (run stale parents in order, then gather, :579-751), ``copy`` functions (:753-852),
* the IPM driver, specialised by the ``#define`` table of
  ``lib/cmex2optimizeCS.m:303-331``.

Two interchangeable sources for the per-instruction C statements and for the driver:

``emitter="ref"``  statements are printed by the REFERENCE's own ``writeCinstructionsC``
                   (``lib/csparse/instructionsTableUTHash.c:790-1609``), loaded from
                   ``oracle/_ref/libitable.so`` (built by ``oracle/build_ref.py``);
``emitter="port"`` statements are printed by :func:`statement`, a restatement of the
                   same switch (forms listed in SURVEY.md Appendix A).
``driver="ref"``   ``#include`` of the unmodified ``/root/reference/lib/ipmPD_CSsolver.c``;
``driver="port"``  ``oracle/ipm_driver_port.c`` (restatement, same control flow).

A small batch runner (``oracle_run_batch``) is appended so that a whole batch can be
solved sequentially with one foreign call per process.
"""
from __future__ import annotations

import ctypes
import os
import tempfile
from typing import Dict, List, Sequence

import numpy as np

from tenscalc_b200 import itable as it
from tenscalc_b200.csparse import CompiledProblem
from tenscalc_b200.optimize import GeneratedSolver

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
MAX_PER_FUNCTION = 200

_DOUBLE_DEFINES = ("gradTolerance", "addEye2HessianUtolerance", "addEye2Hessian1tolerance", "addEye2HessianDtolerance", "equalTolerance", "desiredDualityGap",
                   "alphaMin", "alphaMax", "muFactorAggressive", "muFactorConservative")


# --------------------------------------------------------------------------------------------
# port emitter: one C statement per instruction (instructionsTableUTHash.c:950-1539)
# --------------------------------------------------------------------------------------------
def statement(tab: it.InstructionTable, i: int, slot: np.ndarray) -> str:
    t, p, ops = tab.types[i], tab.params[i], tab.operands[i]
    r = f"m[{slot[i]}]"
    o = [f"m[{slot[k]}]" for k in ops]

    def pairs(xs: Sequence[str]) -> str:
        return "+".join(f"{xs[k]}*{xs[k + 1]}" for k in range(0, len(xs), 2))

    if t == it.I_set:
        return ""
    if t == it.I_load:
        return f"\t{r}={p[0]:.20e};\n"                                        # :955-957
    if t == it.I_sum:
        return f"\t{r}=" + "".join(("+" if s > 0 else "-") + x for s, x in zip(p, o)) + ";\n"   # :964-968
    if t == it.I_sumprod:
        lines = it.sumprod_lines(int(p[0]), o)                                 # :1007-1029 (line breaks :1019-1026)
        return f"\t{r}={lines[0]};\n" + "".join(f"\t{r}+={ln};\n" for ln in lines[1:])
    if t == it.I_plus_minus_dot_div:
        return f"\t{r}=({o[0]}-({pairs(o[2:])}))/{o[1]};\n"                    # :1081-1098
    if t == it.I_minus_dot_div:
        return f"\t{r}=-({pairs(o[1:])})/{o[0]};\n"                            # :1106-1121
    if t == it.I_plus_minus_dot:
        return f"\t{r}={o[0]}-({pairs(o[1:])});\n"                             # :1128-1144
    if t == it.I_minus_dot:
        return f"\t{r}=-({pairs(o)});\n"                                       # :1151-1164
    if t == it.I_div:
        return f"\t{r}={o[0]}/{o[1]};\n"                                       # :1171-1173
    if t == it.I_plus_sqr:
        return f"\t{r}=" + "+".join(f"{x}*{x}" for x in o) + ";\n"             # :1201-1210
    if t == it.I_plus_abs:
        return f"\t{r}=" + "+".join(f"fabs({x})" for x in o) + ";\n"           # :1217-1225
    if t in (it.I_min, it.I_max):
        c = ">" if t == it.I_min else "<"
        return f"\t{r}={o[0]};\n" + "".join(f"\tif ({r}{c}{x}) {r}={x};\n" for x in o[1:])       # :1232-1243
    if t in (it.I_min0, it.I_max0):
        c = ">" if t == it.I_min0 else "<"
        return f"\t{r}=0;\n" + "".join(f"\tif ({r}{c}{x}) {r}={x};\n" for x in o)                # :1250-1259
    if t == it.I_max_abs:
        return f"\t{r}=fabs({o[0]});\n" + "".join(f"\tif ({r}<fabs({x})) {r}=fabs({x});\n" for x in o[1:])  # :1301-1312
    if t == it.I_clp:
        s = f"\t{r}=DBL_MAX;\n"                                                # :1318-1330
        for k in range(0, len(o), 2):
            s += f"\tif ({o[k + 1]}<0) {{SCRATCHBOOK_TYPE x=-{o[k]}/{o[k + 1]}; if ({r}>x) {r}=x;}}\n"
        return s
    x = o[0] if o else ""
    unary = {                                                                  # :1347-1539
        it.I_round: f"round({x})", it.I_ceil: f"ceil({x})", it.I_floor: f"floor({x})",
        it.I_abs: f"fabs({x})", it.I_sign: f"({x}>=0)?1:-1", it.I_2: "2", it.I_2times: f"2*{x}",
        it.I_6times: f"6*{x}", it.I_sqr: f"{x}*{x}", it.I_3sqr: f"3*{x}*{x}", it.I_cube: f"{x}*{x}*{x}",
        it.I_sqrt: f"sqrt({x})", it.I_Dsqrt: f".5*pow({x},-.5)", it.I_DDsqrt: f"-.25*pow({x},-1.5)",
        it.I_inv: f"1/{x}", it.I_minus_inv_sqr: f"-1/({x}*{x})", it.I_2_inv_cube: f"2/({x}*{x}*{x})",
        it.I_cos: f"cos({x})", it.I_minus_cos: f"-cos({x})", it.I_sin: f"sin({x})",
        it.I_minus_sin: f"-sin({x})", it.I_log: f"log({x})", it.I_exp: f"exp({x})", it.I_atan: f"atan({x})",
        it.I_Datan: f"1/(1+{x}*{x})", it.I_DDatan: f"-2*{x}/(1+{x}*{x})/(1+{x}*{x})",
        it.I_srelu: f"log(1+exp({x}))", it.I_dsrelu: f"1/(1+exp(-{x}))",
        it.I_ddsrelu: f"1/(2+exp(-{x})+exp({x}))", it.I_relu: f"({x}>0)?{x}:0",
        it.I_heaviside: f"({x}>0)?1:0",
    }
    if t == it.I_componentwise:                                               # :1333-1345
        return f"\t{r}=" + it.componentwise_text(p, x) + ";\n"
    if t in unary:
        return f"\t{r}={unary[t]};\n"
    raise NotImplementedError(it.OPNAMES.get(t, str(t)))


# --------------------------------------------------------------------------------------------
# reference emitter through ctypes
# --------------------------------------------------------------------------------------------
class RefEmitter:
    """Feeds the instruction table to the reference's store and lets its
    ``writeCinstructionsC`` print the statements of each group."""

    def __init__(self, problem: CompiledProblem):
        path = os.path.join(REF_DIR, "libitable.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `python oracle/build_ref.py` (needs /root/reference)")
        lib = ctypes.CDLL(path)
        lib.appendInstruction.restype = ctypes.c_int64
        lib.appendInstruction.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.POINTER(ctypes.c_double),
                                          ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
        lib.writeCinstructionsC.restype = ctypes.c_int
        lib.writeCinstructionsC.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int64),
                                            ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                            ctypes.POINTER(ctypes.c_int64), ctypes.c_size_t, ctypes.c_size_t,
                                            ctypes.c_size_t]
        lib.initInstructionsTable()
        tab = problem.tab
        # only live instructions are appended; ref index (1-based) per table id
        self.ref_index = np.zeros(len(tab), dtype=np.int64)
        live = np.nonzero(problem.slot >= 0)[0]
        for i in live:
            p = np.asarray(tab.params[i], dtype=np.float64)
            o = np.asarray([self.ref_index[k] for k in tab.operands[i]], dtype=np.int64)
            assert (o > 0).all()
            idx = lib.appendInstruction(int(tab.types[i]), len(p), p.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                        len(o), o.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
            assert idx > 0
            self.ref_index[i] = idx
        self.n_ref = int(self.ref_index.max())
        self.mem = np.zeros(self.n_ref, dtype=np.int64)          # 1-based memory locations
        self.mem[self.ref_index[live] - 1] = problem.slot[live] + 1
        self.lib = lib
        self.flops = np.zeros(17, dtype=np.int64)

    def statements(self, ids: Sequence[int], min4loop: int = 50) -> str:
        idx = np.asarray([self.ref_index[i] for i in ids], dtype=np.int64)
        m4l = ctypes.c_int64(min4loop)
        with tempfile.NamedTemporaryFile("r", suffix=".c", delete=True) as f:
            name = f.name.encode()
            rc = self.lib.writeCinstructionsC(name, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                              self.mem.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                              ctypes.byref(m4l),
                                              self.flops.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                              len(name), len(idx), self.n_ref)
            assert rc == 0
            return open(f.name).read()


# --------------------------------------------------------------------------------------------
# whole translation unit
# --------------------------------------------------------------------------------------------
def _run_groups(groups: Sequence[int]) -> str:
    return "".join(f"  if (groupStatus[{g}]==0) {{groupStatus[{g}]=1;group{g}();}}\n" for g in groups)


def _invalidate(groups: Sequence[int]) -> str:
    return "".join(f"  groupStatus[{g}]=0;\n" for g in groups)


def emit_translation_unit(gen: GeneratedSolver, emitter: str = "port", driver: str = "port") -> str:
    pr = gen.problem
    tab, slot = pr.tab, pr.slot
    ref = RefEmitter(pr) if emitter == "ref" else None
    out: List[str] = []
    w = out.append
    w("/* generated by oracle/emit_c.py -- CPU oracle, test infrastructure only */\n")
    w("#include <math.h>\n#include <float.h>\n#include <string.h>\n#include <stdint.h>\n#include <stdlib.h>\n")
    w("#define SCRATCHBOOK_TYPE double\n#define EXPORT __attribute__((visibility(\"default\")))\n")
    # libm-sensitivity probe (Oracle(..., variant="crpow"), built with -DORACLE_CRPOW): the CORRECTLY ROUNDED
    # x^-1/2 and x^-3/2 in place of glibc's pow, which differs from them on ~8e-4 of arguments.  Comparing the
    # two oracles on the same batch measures how many instances change iteration count when ONE libm result in
    # a thousand moves by one ULP -- the noise floor of any parity claim on problems that call pow.
    w("#ifdef ORACLE_CRPOW\n"
      "static void oracle_rsqrt_dd(double x,double *hi,double *lo){double r=1.0/sqrt(x),t=r*r,tl=fma(r,r,-t),s=x*t,\n"
      "  sl=fma(x,t,-s)+x*tl,e=(1.0-s)-sl; *hi=r; *lo=0.5*r*e;}\n"
      "static double oracle_crpow(double x,double y){double h,l;\n"
      "  if (y==-.5 && x>1e-290 && x<1e290){oracle_rsqrt_dd(x,&h,&l); return h+l;}\n"
      "  if (y==-1.5 && x>1e-190 && x<1e190){oracle_rsqrt_dd(x,&h,&l); double q0=h/x,rem=fma(-q0,x,h); return q0+(rem+l)/x;}\n"
      "  return pow(x,y);}\n"
      "#define pow(x,y) oracle_crpow(x,y)\n#endif\n")
    solver = str(pr.defines.get("driver", "ipmPD_CSsolver"))       # which of the reference's drivers the class calls
    for k, v in pr.defines.items():
        if k != "driver":
            w(f"#define {k} {v:e}\n" if k in _DOUBLE_DEFINES else f"#define {k} {v}\n")
    w(f"static SCRATCHBOOK_TYPE scratchbook[{max(pr.n_slots, 1)}];\n")
    w(f"static int groupStatus[{max(len(pr.groups), 1)}];\n")
    # variant "count": executions of every dependency group -- the counters behind the reference's profiling option
    # (writeCfunctionpergroup.m:149-177,481-492), which examples/flops.m:94-99 turns into flops per solve
    w(f"#ifdef ORACLE_COUNT\nstatic long long oracle_group_count[{max(len(pr.groups), 1)}];\n"
      f"EXPORT void oracle_group_counts(long long *out, int reset) {{ for (int g=0; g<{len(pr.groups)}; g++) "
      "{ out[g]=oracle_group_count[g]; if (reset) oracle_group_count[g]=0; } }\n#endif\n")
    for g, members in enumerate(pr.groups):
        chunks = [members[k:k + MAX_PER_FUNCTION] for k in range(0, len(members), MAX_PER_FUNCTION)]
        for c, chunk in enumerate(chunks):
            body = ref.statements(chunk) if ref else "".join(statement(tab, i, slot) for i in chunk)
            w(f"static void group{g}_{c}(void) {{\n  SCRATCHBOOK_TYPE *m=scratchbook;\n{body}}}\n")
        w(f"static void group{g}(void) {{\n#ifdef ORACLE_COUNT\n  oracle_group_count[{g}]++;\n#endif\n"
          + "".join(f"  group{g}_{c}();\n" for c in range(len(chunks))) + "}\n")

    for s in pr.sets.values():
        fname = f"{gen.classname}_set_{s.name[5:]}" if s.kind != "solver" else s.name
        w(f"EXPORT void {fname}(const double *input) {{\n  SCRATCHBOOK_TYPE *m=scratchbook;\n")
        w("".join(f"  m[{slot[d]}]=*(input++);\n" for d in s.dest))
        w(_invalidate(s.children) + "}\n")
    if int(pr.defines.get("allowSave", 0)):
        # allowSave: the reference's driver calls saveWW__(saveNamePrefix "_WW.values") (ipmPD_CSsolver.c:51,330-332,
        # 561-566,671-676,699-703,877-882).  The oracle's saveWW__ is the copy of the KKT values into the save
        # variables that getOutputs returns (optimize.py), the file name is ignored (the reference's driver
        # still prints its "Saving ..." line per event; the port driver has the same hooks, without the print).
        w('#define saveNamePrefix "oracle"\n')
    for c in pr.copies.values():
        sig = "char *filename" if c.name == "saveWW__" else "void"
        w(f"EXPORT void {c.name}({sig}) {{\n  SCRATCHBOOK_TYPE *m=scratchbook;\n" + _run_groups(c.parents))
        w("".join(f"  m[{slot[d]}]=m[{slot[s]}];\n" for d, s in zip(c.dest, c.src) if d != s))
        w(_invalidate(c.children) + "}\n")
    for g in pr.gets.values():
        fname = f"{gen.classname}_getOutputs" if g.name == "getOutputs" else g.name
        args = ",".join(f"double *y{k + 1}" for k in range(len(g.sources)))
        w(f"EXPORT void {fname}({args}) {{\n  SCRATCHBOOK_TYPE *m=scratchbook;\n" + _run_groups(g.parents))
        for k, src in enumerate(g.sources):
            w("".join(f"  y{k + 1}[{j}]=" + ("0" if x is None else f"m[{slot[x]}]") + ";\n"
                      for j, x in enumerate(src)))
        w("}\n")
    # functions the reference driver references under other option sets
    for name, sig in (("getHessInertia__", "double *mp,double *mn"),):
        if name not in pr.gets:
            w(f"static void {name}({sig}) {{ *mp=0; *mn=0; }}\n")

    if solver == "ipmPDminmax_CSsolver":    # PARITY UNPINNED: transcription of the reference's MATLAB solver (see that file)
        w(f'#include "{os.path.join(HERE, "ipm_driver_minmax_port.c")}"\n')
    elif driver == "ref":
        w(f'#include "/root/reference/lib/{solver}.c"\n')
    elif solver != "ipmPD_CSsolver":
        raise NotImplementedError(f"{solver}: only the reference's own driver is used as oracle (kind='ref')")
    else:
        w(f'#include "{os.path.join(HERE, "ipm_driver_port.c")}"\n')

    # ---- batch runner ------------------------------------------------------------------------
    npar = [(f"{gen.classname}_set_{n}", int(np.prod(s)) if s else 1) for n, s in gen.parameters]
    nvar = [(f"{gen.classname}_set_{n}", int(np.prod(s)) if s else 1) for n, s in gen.variables]
    nout = [int(np.prod(s)) if s else 1 for _, s in gen.outputs]
    w("#include <time.h>\n")
    w("EXPORT void oracle_run_batch(int64_t B,const double *P,const double *V,double mu0,int32_t maxIter,\n"
      "    int32_t saveIter,const double *addEye,int32_t *status,int32_t *iter,double *Y,double *seconds) {\n"
      f"  const int64_t nP={sum(n for _, n in npar)},nV={sum(n for _, n in nvar)},nY={sum(nout)};\n"
      "  for (int64_t b=0;b<B;b++) {\n    const double *p=P+b*nP,*v=V+b*nV; double *y=Y+b*nY;\n"
      "    double mu=mu0,ae[3]={addEye[0],addEye[1],addEye[2]},t; int32_t mi=maxIter,si=saveIter;\n")
    off = 0
    for f, n in npar:
        w(f"    {f}(p+{off});\n")
        off += n
    off = 0
    for f, n in nvar:
        w(f"    {f}(v+{off});\n")
        off += n
    w("    struct timespec t0,t1; clock_gettime(CLOCK_MONOTONIC,&t0);\n")
    w(f"    status[b]=0; {solver}(&mu,&mi,&si,ae,status+b,iter+b,&t);\n")
    w("    clock_gettime(CLOCK_MONOTONIC,&t1);\n")
    w("    if (seconds) seconds[b]=(t1.tv_sec-t0.tv_sec)+1e-9*(t1.tv_nsec-t0.tv_nsec);\n")
    offs = np.concatenate([[0], np.cumsum(nout)])
    w(f"    {gen.classname}_getOutputs(" + ",".join(f"y+{offs[k]}" for k in range(len(nout))) + ");\n")
    w("  }\n}\n")
    return "".join(out)
