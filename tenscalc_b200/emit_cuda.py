"""sm_100a CUDA emitter: the straight-line code of every dependency group as CUDA C.

Counterpart of the reference's C emitter (``writeCinstructionsC``,
``lib/csparse/instructionsTableUTHash.c:790-1609`` driven by
``lib/@csparse/writeCfunctionpergroup.m:410-443``), re-designed for one-instance-per-lane
execution on Blackwell:

* every group becomes a sequence of CHUNKS; ``merge`` consecutive chunks share one function
  ``tcg<g>_<k>(double *m, unsigned mask)``.  ``m`` points at the lane's column of the warp's
  scratchbook tile, so slot ``s`` of the instance is ``m[32*s]`` and a warp-wide access is one
  contiguous 256-byte segment (``tc_image.h``); ``mask`` says for which of the lane's instances
  (``tiles`` per lane, normally one) the code runs;
* a chunk loads ALL its scratchbook operands up front, back to back (at most ``max_loads``
  distinct ones: they stay in registers), so it pays one memory round trip, and while they are
  in flight it prefetches the operands of the NEXT chunk into L2 (``prefetch``);
* inside a chunk every instruction result is a named local (a register) and is written back to
  the scratchbook only if somebody outside the chunk reads it (another chunk, another group, a
  get or a copy) -- short-lived temporaries never touch HBM;
* constants are immediates;
* the expression of every instruction is the C expression of the reference emitter,
  evaluated in the same order; the translation units are compiled with ``-fmad=false`` so
  multiplies and adds round separately like the reference's x86-64 build.

``emit_module`` spreads the functions over part files of equal size (compiled in parallel with
``-rdc``, objects cached by content) and writes ``main.cu``: the group/copy dispatchers, the
device-resident solver loop (``csrc/tc_driver.cuh`` or, ``driver="lanes"``,
``csrc/tc_driver_lanes.cuh``), the kernels and the ``tcm_*`` entry points that
``tc_attach_module`` binds.  Options and what they measured to on B200 are listed in DESIGN.md
section 3.1; the defaults are the fastest measured configuration.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Set, Tuple

from . import itable as it
from .csparse import CompiledProblem

TILE = 32                  # lanes per scratchbook tile -- MUST match TC_TILE


def _lit(v: float) -> str:
    if v != v or v in (float("inf"), float("-inf")):
        raise ValueError("non-finite constant")
    return float(v).hex() if v != int(v) or abs(v) >= 1e15 else f"{int(v)}.0"


def expression(t: int, p: Sequence[float], o: List[str], r: str) -> str:
    """CUDA statements computing instruction (t,p) with operand expressions ``o`` into the
    new local ``r``; forms and order of instructionsTableUTHash.c:950-1539."""

    def pairs(xs: Sequence[str]) -> str:
        return "+".join(f"{xs[k]}*{xs[k + 1]}" for k in range(0, len(xs), 2))

    if t == it.I_sum:                                                         # :964-968
        return f"const double {r}=" + "".join(("+" if s > 0 else "-") + x for s, x in zip(p, o)) + ";"
    if t == it.I_sumprod:                                                     # :1007-1029
        lines = it.sumprod_lines(int(p[0]), o)
        if len(lines) == 1:
            return f"const double {r}={lines[0]};"
        # more than MAX_TERMS_PERLINE factors: the reference breaks the statement, m[d]=line1; m[d]+=line2; ...
        # -- every line is summed on its own and then added to the running value (:1019-1026)
        return f"double {r}={lines[0]};" + "".join(f" {r}+={ln};" for ln in lines[1:])
    if t == it.I_plus_minus_dot_div:                                          # :1081-1098
        return f"const double {r}=TC_DIV(({o[0]}-({pairs(o[2:])})),{o[1]});"
    if t == it.I_minus_dot_div:                                               # :1106-1121
        return f"const double {r}=TC_DIV(-({pairs(o[1:])}),{o[0]});"
    if t == it.I_plus_minus_dot:                                              # :1128-1144
        return f"const double {r}={o[0]}-({pairs(o[1:])});"
    if t == it.I_minus_dot:                                                   # :1151-1164
        return f"const double {r}=-({pairs(o)});"
    if t == it.I_div:
        return f"const double {r}=TC_DIV({o[0]},{o[1]});"
    if t == it.I_plus_sqr:                                                    # :1201-1210
        return f"const double {r}=" + "+".join(f"{x}*{x}" for x in o) + ";"
    if t == it.I_plus_abs:                                                    # :1217-1225
        return f"const double {r}=" + "+".join(f"fabs({x})" for x in o) + ";"
    if t in (it.I_min, it.I_max):                                             # :1232-1243
        c = ">" if t == it.I_min else "<"
        return f"double {r}={o[0]};" + "".join(f" if ({r}{c}{x}) {r}={x};" for x in o[1:])
    if t in (it.I_min0, it.I_max0):                                           # :1250-1259
        c = ">" if t == it.I_min0 else "<"
        return f"double {r}=0.0;" + "".join(f" if ({r}{c}{x}) {r}={x};" for x in o)
    if t == it.I_max_abs:                                                     # :1301-1312
        return f"double {r}=fabs({o[0]});" + "".join(f" if ({r}<fabs({x})) {r}=fabs({x});" for x in o[1:])
    if t == it.I_clp:                                                         # :1318-1330
        s = f"double {r}=DBL_MAX;"
        for k in range(0, len(o), 2):
            s += f" if ({o[k + 1]}<0) {{const double x_=-{o[k]}/{o[k + 1]}; if ({r}>x_) {r}=x_;}}"
        return s
    if t == it.I_componentwise:                                               # :1333-1345
        # the user's C text with the operand at every placeholder; on the device it is compiled as CUDA C, so it
        # may call what <cmath> offers there (the text is the user's: its arithmetic is theirs)
        return f"const double {r}=" + it.componentwise_text(p, f"({o[0]})") + ";"
    x = o[0] if o else ""
    unary = {                                                                 # :1347-1539
        it.I_round: f"round({x})", it.I_ceil: f"ceil({x})", it.I_floor: f"floor({x})",
        it.I_abs: f"fabs({x})", it.I_sign: f"({x}>=0)?1.0:-1.0", it.I_2times: f"2.0*{x}",
        it.I_6times: f"6.0*{x}", it.I_sqr: f"{x}*{x}", it.I_3sqr: f"3.0*{x}*{x}", it.I_cube: f"{x}*{x}*{x}",
        it.I_sqrt: f"sqrt({x})", it.I_Dsqrt: f".5*tc_pow_m05({x})", it.I_DDsqrt: f"-.25*tc_pow_m15({x})",
        it.I_inv: f"1.0/{x}", it.I_minus_inv_sqr: f"-1.0/({x}*{x})", it.I_2_inv_cube: f"2.0/({x}*{x}*{x})",
        it.I_cos: f"cos({x})", it.I_minus_cos: f"-cos({x})", it.I_sin: f"sin({x})",
        it.I_minus_sin: f"-sin({x})", it.I_log: f"log({x})", it.I_exp: f"exp({x})", it.I_atan: f"atan({x})",
        it.I_Datan: f"1.0/(1.0+{x}*{x})", it.I_DDatan: f"-2.0*{x}/(1.0+{x}*{x})/(1.0+{x}*{x})",
        it.I_srelu: f"log(1.0+exp({x}))", it.I_dsrelu: f"1.0/(1.0+exp(-{x}))",
        it.I_ddsrelu: f"1.0/(2.0+exp(-{x})+exp({x}))", it.I_relu: f"({x}>0)?{x}:0.0",
        it.I_heaviside: f"({x}>0)?1.0:0.0",
    }
    if t in unary:
        return f"const double {r}={unary[t]};"
    raise NotImplementedError(it.OPNAMES.get(t, str(t)))


def external_values(pr: CompiledProblem) -> Set[int]:
    """values that must reach the scratchbook whatever the chunking: read by another group, a
    get or a copy"""
    tab = pr.tab
    ext: Set[int] = set()
    for g in pr.gets.values():
        ext.update(x for s in g.sources for x in s if x is not None)
    for c in pr.copies.values():
        ext.update(c.src)
    for i, gi in pr.group_of.items():
        for o in tab.operands[i]:
            if pr.group_of.get(o, -1) != gi:
                ext.add(o)
    return ext


def chunk_group(pr: CompiledProblem, members: Sequence[int], max_refs: int, max_loads: int) -> List[List[int]]:
    """split a group into chunks: at most ``max_refs`` operand references and at most ``max_loads``
    distinct scratchbook operands per chunk (they are all loaded up front and stay in registers)"""
    tab = pr.tab
    chunks: List[List[int]] = []
    cur: List[int] = []
    refs = 0
    loaded: Set[int] = set()
    made: Set[int] = set()
    for i in members:
        ops = [o for o in tab.operands[i] if tab.types[o] not in (it.I_load, it.I_2)]
        new = {o for o in ops if o not in loaded and o not in made}
        if cur and (refs + len(ops) > max_refs or len(loaded) + len(new) > max_loads):
            chunks.append(cur)
            cur, refs, loaded, made = [], 0, set(), set()
            new = set(ops)
        cur.append(i)
        refs += len(ops)
        loaded |= new
        made.add(i)
    if cur:
        chunks.append(cur)
    return chunks


ROLLABLE = (it.I_sum, it.I_sumprod, it.I_div)      # the types writeCinstructionsC emits as for(;;) loops
MAX_DELTA_OPERANDS = 10                            # instructionsTableUTHash.c:784


def rolled_runs(pr: CompiledProblem, members: Sequence[int], min_len: int) -> List[Tuple[int, int]]:
    """[(position in ``members``, count)] of the runs the reference emitter rolls into ``for`` loops
    (instructionsTableUTHash.c:881-944): consecutive instructions of equal type, parameter vector and
    operand count (at most 10 operands) whose result slot and every operand slot advance by constant
    strides; scanned greedily from every instruction, rolled when the run holds at least ``min_len``
    (``minInstructions4loop``, default 50) instructions of type sum, sumprod or div."""
    tab, slot = pr.tab, pr.slot
    n = len(members)

    def same_shape(a: int, b: int) -> bool:
        return (tab.types[a] == tab.types[b] and len(tab.operands[a]) == len(tab.operands[b])
                and list(tab.params[a]) == list(tab.params[b]))

    runs: List[Tuple[int, int]] = []
    i = 0
    while i < n:
        a = members[i]
        nxt = i + 1
        if nxt < n and len(tab.operands[a]) <= MAX_DELTA_OPERANDS and same_shape(a, members[nxt]):
            b = members[nxt]
            d_res = int(slot[b]) - int(slot[a])
            d_ops = [int(slot[y]) - int(slot[x]) for x, y in zip(tab.operands[a], tab.operands[b])]
            nxt += 1
            while nxt < n:
                prev, c = members[nxt - 1], members[nxt]
                if int(slot[c]) - int(slot[prev]) != d_res or not same_shape(a, c):
                    break
                if [int(slot[y]) - int(slot[x]) for x, y in zip(tab.operands[prev], tab.operands[c])] != d_ops:
                    break
                nxt += 1
        if tab.types[a] in ROLLABLE and nxt - i >= min_len:
            runs.append((i, nxt - i))
            i = nxt
        else:
            i += 1
    return runs


def _device_rollable(pr: CompiledProblem, run: Sequence[int]) -> bool:
    """a run becomes a device loop when every operand position is either the SAME constant in all its
    instructions (an immediate) or a scratchbook value in all of them (constants live in no scratchbook)"""
    tab = pr.tab
    const_types = (it.I_load, it.I_2)
    first = tab.operands[run[0]]
    for i in run[1:]:
        for x, y in zip(first, tab.operands[i]):
            if (tab.types[x] in const_types or tab.types[y] in const_types) and x != y:
                return False
    return True


def chunk_group_rolled(pr: CompiledProblem, members: Sequence[int], max_refs: int, max_loads: int,
                       roll: int) -> Tuple[List[List[int]], Set[int]]:
    """``chunk_group`` with every rolled run (``rolled_runs`` of at least ``roll`` instructions) as a chunk of its
    own: (chunks, indices of the rolled chunks)"""
    chunks: List[List[int]] = []
    rolled: Set[int] = set()
    pos = 0
    for start, count in rolled_runs(pr, members, roll) if roll else []:
        run = list(members[start:start + count])
        if not _device_rollable(pr, run):
            continue
        if start > pos:
            chunks += chunk_group(pr, members[pos:start], max_refs, max_loads)
        rolled.add(len(chunks))
        chunks.append(run)
        pos = start + count
    if pos < len(members) or not chunks:
        chunks += chunk_group(pr, members[pos:], max_refs, max_loads)
    return chunks, rolled


def emit_rolled(pr: CompiledProblem, run: Sequence[int], W: int) -> Tuple[List[str], int]:
    """statements of one rolled run: ``for`` over the instructions of the run, slot of every operand and of the
    result = first slot + counter x stride.  Runs without a loop-carried value (no instruction of the run reads
    a slot the run writes) fetch the operands of ``U`` iterations back to back before computing them (one memory
    round trip per ``U`` instructions, like a chunk); the others run strictly in the reference's order."""
    tab, slot = pr.tab, pr.slot
    const_types = (it.I_load, it.I_2)
    a, n = run[0], len(run)
    t, p = tab.types[a], tab.params[a]
    ops = tab.operands[a]
    d_res = int(slot[run[1]]) - int(slot[a])
    d_ops = [int(slot[y]) - int(slot[x]) for x, y in zip(ops, tab.operands[run[1]])]
    var = [j for j, q in enumerate(ops) if tab.types[q] not in const_types]
    written = {int(slot[i]) for i in run}
    carried = any(int(slot[q]) in written for i in run for q in tab.operands[i] if tab.types[q] not in const_types)
    U = 1 if carried else max(1, min(8, 48 // max(len(var), 1)))

    def addr(s0: int, d: int) -> str:
        return f"mt[({s0}+(c0+u)*({d}))*{TILE * W}]"

    o: List[str] = []
    for j, q in enumerate(ops):
        if tab.types[q] == it.I_load:
            o.append("(" + _lit(tab.params[q][0]) + ")")
        elif tab.types[q] == it.I_2:
            o.append("2.0")
        else:
            o.append(f"x_[u][{var.index(j)}]")
    lines = [f"// rolled run: {n} x {it.OPNAMES.get(t, t)} (instructionsTableUTHash.c:881-944)"
             + (", loop-carried" if carried else ""),
             f'_Pragma("unroll 1") for (int c0 = 0; c0 < {n}; c0 += {U}) {{',
             f"double x_[{U}][{max(len(var), 1)}];",
             f'_Pragma("unroll") for (int u = 0; u < {U}; u++) if (c0 + u < {n}) {{']
    lines += [f"x_[u][{k}]={addr(int(slot[ops[j]]), d_ops[j])};" for k, j in enumerate(var)]
    lines += ["}", f'_Pragma("unroll") for (int u = 0; u < {U}; u++) if (c0 + u < {n}) {{',
              expression(t, p, o, "r_"), f"{addr(int(slot[a]), d_res)}=r_;", "}", "}"]
    return lines, len(var)


def plan_scache(reads: List[List[int]], makes: List[List[int]], ext: Set[int], capacity: int):
    """Offline allocation of a per-lane shared-memory cache of ``capacity`` doubles for one group.

    The access trace of a group is static, so the replacement policy can be the optimal one (Belady:
    after every chunk keep the ``capacity`` values whose next use is nearest) and every decision is taken
    here, at generation time: the emitted code holds plain shared-memory reads and writes at fixed
    slots, no tags, no lookups.  ``reads[k]`` are the distinct scratchbook operands chunk k takes from
    outside itself, ``makes[k]`` the values it produces.  Returns per chunk
    ``(from_smem {value: slot}, to_smem [(value, slot)])`` and the set of produced values that must
    ALSO be stored in the scratchbook (read outside the group, or dropped from the cache before their
    last use); a produced value that stays cached until its last use never touches HBM."""
    INF = 1 << 60
    uses: Dict[int, List[int]] = {}
    for k, rs in enumerate(reads):
        for q in rs:
            uses.setdefault(q, []).append(k)
    ptr: Dict[int, int] = {}

    def next_use(q: int, k: int) -> int:
        u = uses.get(q)
        if not u:
            return INF
        p = ptr.get(q, 0)
        while p < len(u) and u[p] <= k:
            p += 1
        ptr[q] = p
        return u[p] if p < len(u) else INF

    cache: Dict[int, int] = {}             # value -> slot
    free = list(range(capacity - 1, -1, -1))
    plan = []
    to_hbm: Set[int] = set()
    made_all = {i for ms in makes for i in ms}
    for k, (rs, ms) in enumerate(zip(reads, makes)):
        from_smem = {q: cache[q] for q in rs if q in cache}
        cand = [(next_use(q, k), n, q) for n, q in enumerate(list(cache) + [q for q in rs if q not in cache] + list(ms))]
        cand = sorted(c for c in cand if c[0] < INF)
        keep = {q for _, _, q in cand[:capacity]}
        for q in [q for q in cache if q not in keep]:
            free.append(cache.pop(q))
        for _, _, q in cand[capacity:]:
            if q in made_all:              # leaves (or never enters) the cache with uses ahead: needs its HBM copy
                to_hbm.add(q)
        to_smem = []
        for _, _, q in cand[:capacity]:
            if q not in cache:
                cache[q] = free.pop()
                to_smem.append((q, cache[q]))
        plan.append((from_smem, to_smem))
    to_hbm |= made_all & ext
    return plan, to_hbm


def emit_group(pr: CompiledProblem, g: int, ext: Set[int], max_refs: int = 1536,
               max_loads: int = 96, prefetch: int = 1, merge: int = 1,
               assume_global: int = 0, barrier: int = 0, scache: int = 0,
               pair: int = 0, wide: int = 0, roll: int = 0) -> Tuple[List[Tuple[str, str]], Dict[str, int]]:
    """[(function name, CUDA source)] of the chunks of group g, plus traffic statistics.

    ``roll`` > 0: runs of at least ``roll`` strided repeats of one sum / sumprod / div instruction become
    device loops (``rolled_runs``: the reference's loop rolling, ``minInstructions4loop``), each a chunk of its
    own that reads and writes the scratchbook directly.

    ``pair``: every lane carries TWO instances whose scratchbook columns are interleaved
    (``m[tile pair][slot][lane][2]``): a chunk runs the group on both -- one 128-bit load or store
    moves slot s of both instances, the arithmetic is emitted twice (two independent dependency
    chains per lane).  The chunk runs on both instances whenever either needs the group:
    re-evaluating a group on unchanged inputs rewrites the same values, evaluating one whose inputs
    are stale only touches values that are recomputed before anybody reads them (the caller marks
    the group valid only for the instances that asked)."""
    W = 2 if pair or wide else 1   # slot stride in tiles of 32 doubles (pair: two doubles per lane; wide: 64-lane tiles)
    ty = "double2" if pair else "double"
    insts = ("x", "y") if pair else ("",)
    tab, slot = pr.tab, pr.slot
    members = pr.groups[g]
    chunks, rolled = chunk_group_rolled(pr, members, max_refs, max_loads, roll)
    chunk_of = {i: k for k, ch in enumerate(chunks) for i in ch}
    const_types = (it.I_load, it.I_2)
    skip_types = (it.I_set, it.I_load, it.I_2)
    # results read by a later chunk of the same group
    cross: Set[int] = set()
    for i in members:
        for o in tab.operands[i]:
            k = chunk_of.get(o)
            if k is not None and k != chunk_of[i]:
                cross.add(o)
    # per chunk: the distinct scratchbook operands it takes from outside itself (in order of first use) and
    # the values it produces
    reads: List[List[int]] = []
    makes: List[List[int]] = []
    for ch in chunks:
        have: Set[int] = set()
        rs: List[int] = []
        ms: List[int] = []
        for i in ch:
            if tab.types[i] in skip_types:
                continue
            for q in tab.operands[i]:
                if tab.types[q] not in const_types and q not in have:
                    have.add(q)
                    rs.append(q)
            have.add(i)
            ms.append(i)
        reads.append(rs)
        makes.append(ms)
    if scache:
        plan, to_hbm = plan_scache(reads, makes, ext, scache)
    else:
        plan, to_hbm = [({}, []) for _ in chunks], ext | cross
    stats = dict(loads=0, stores=0, reg_reads=0, reg_only=0, prefetches=0, smem_loads=0, smem_stores=0, rolled=0)
    bodies = []                    # per chunk: (name, HBM load slots in issue order, store slots, statements)
    def ld(q: int, e: str) -> str:          # a fetched operand of instance e: the vector's component
        return f"v{q}.{e}" if pair else f"v{q}"

    def made(q: int, e: str) -> str:        # a result of instance e
        return f"v{q}{e}"

    for k, ch in enumerate(chunks):
        name = f"tcg{g}_{k}"
        if k in rolled:                       # a loop over the run: operands from / results to the scratchbook
            body, nvar = emit_rolled(pr, ch, W)
            stats["loads"] += nvar * len(ch)
            stats["stores"] += len(ch)
            stats["rolled"] += len(ch)
            bodies.append((name, [], {int(slot[i]) for i in ch}, [], body))
            continue
        from_smem, to_smem = plan[k]
        keep_slot = dict(to_smem)
        fetched = set(reads[k])
        # every operand of the chunk is fetched up front, back to back: one memory round trip per chunk instead of
        # one per use (the scratchbook loads first, then the shared-memory cache hits)
        if pair:
            head = [f"const double2 v{q}=*(const double2 *)(sc+{from_smem[q] * TILE * W});" for q in reads[k] if q in from_smem]
        else:
            head = [f"const double v{q}=sc[{from_smem[q] * TILE}];" for q in reads[k] if q in from_smem]   # (the cache is per warp)
        loads = [q for q in reads[k] if q not in from_smem]
        lines: List[str] = []
        tail: List[str] = []       # loaded operands that enter the cache: written at the end, when they have arrived
        stores: Set[int] = set()
        stats["loads"] += len(loads)
        stats["smem_loads"] += len(head)
        for q in loads:
            if q in keep_slot:
                tail.append(f"*(double2 *)(sc+{keep_slot[q] * TILE * W})=v{q};" if pair else f"sc[{keep_slot[q] * TILE}]=v{q};")
                stats["smem_stores"] += 1
        for i in ch:
            t, p = tab.types[i], tab.params[i]
            if t in skip_types:               # constants are immediates at their uses (gets/copies read the pool)
                continue
            for e in insts:
                o: List[str] = []
                for q in tab.operands[i]:
                    tq = tab.types[q]
                    if tq == it.I_load:
                        o.append("(" + _lit(tab.params[q][0]) + ")")
                    elif tq == it.I_2:
                        o.append("2.0")
                    else:
                        o.append(ld(q, e) if q in fetched else made(q, e))
                lines.append(expression(t, p, o, made(i, e)))
            stats["reg_reads"] += len(tab.operands[i])
            both = f"make_double2(v{i}x,v{i}y)" if pair else f"v{i}"
            if i in keep_slot:
                lines.append(f"*(double2 *)(sc+{keep_slot[i] * TILE * W})={both};" if pair else f"sc[{keep_slot[i] * TILE}]={both};")
                stats["smem_stores"] += 1
            if i in to_hbm:
                lines.append(f"*(double2 *)(mt+{int(slot[i]) * TILE * W})={both};" if pair else f"mt[{int(slot[i]) * TILE * W}]={both};")
                stores.add(int(slot[i]))
                stats["stores"] += 1
            elif i not in keep_slot:
                stats["reg_only"] += 1
        bodies.append((name, loads, stores, head, lines + tail))
    out: List[Tuple[str, str]] = []
    # `merge` consecutive chunks share one function (each chunk in its own scope): the ABI's
    # callee-saved register spill/restore is paid once per function instead of once per chunk
    for f0 in range(0, len(bodies), merge):
        fname = bodies[f0][0]
        src = [f"__device__ __noinline__ void {fname}(double *__restrict__ m, unsigned mask) {{"]
        if assume_global:          # LDG/STG instead of generic LD/ST (the pointer's address space is not
            src.append("__builtin_assume(__isGlobal(m));")   # visible across -rdc translation units)
        if scache:                 # the lane's column of its warp's cache: [slot][lane], conflict-free
            src.append("double *__restrict__ sc = TC_SCACHE;")   # (pair: 16 bytes per lane and slot)
        for k in range(f0, min(f0 + merge, len(bodies))):
            name, loads, stores, head, lines = bodies[k]
            # one pass over the chunk's code per tile of the warp: the second pass runs out of the
            # instruction cache (TC_TILE_LOOP / TC_MT are defined in the module header)
            # lockstep: the warps of the CTA enter every chunk together, so its code is fetched once per CTA while
            # it is still in the instruction caches.  (Every barrier ends a basic block and ptxas re-creates the
            # uniform registers of the memory descriptors in each -- ~0.5 R2UR per load; a barrier every 4th
            # chunk halves them and measures the same, profiles/r03_experiments.md.)
            if barrier:
                src.append("__syncthreads();")
            src.append("if (mask) {" if pair else "TC_TILE_LOOP {")
            src.append("double *__restrict__ mt = m;" if pair else "double *__restrict__ mt = TC_MT;")
            if pair:
                src += [f"const double2 v{q}=*(const double2 *)(mt+{int(slot[q]) * TILE * W});" for q in loads]
            else:
                src += [f"const double v{q}=mt[{int(slot[q]) * TILE * W}];" for q in loads]
            src += head
            # while this chunk's operands are in flight, pull the operands of the chunk `prefetch` positions ahead
            # into L2: its load phase then costs an L2 hit instead of an HBM round trip (no registers involved);
            # with a distance above 1 the first chunks of the group also cover the chunks in between
            targets = [k + prefetch] if prefetch else []
            if prefetch > 1 and k == 0:
                targets = list(range(1, prefetch + 1))
            nxt: List[int] = []
            mine = {int(slot[q]) for q in loads} | stores
            for kt in targets:
                if kt < len(bodies):
                    between = set().union(*[bodies[j][2] for j in range(k + 1, kt)]) if kt > k + 1 else set()
                    nxt += [int(slot[q]) for q in bodies[kt][1] if int(slot[q]) not in mine and int(slot[q]) not in between]
            if nxt:                # the clobber keeps this chunk's own loads ahead of the prefetches
                src.append('asm volatile("" ::: "memory");')
                src += [f"TC_PREFETCH(mt+{sl * TILE * W});" for sl in dict.fromkeys(nxt)]
                stats["prefetches"] += len(set(nxt))
            src += lines
            src.append("}")
        src.append("}")
        out.append((fname, "\n".join(src) + "\n"))
    return out, stats


# --------------------------------------------------------------------------------------------
# per-problem module: chunk functions spread over part files + main.cu (dispatch, kernels, ABI)
# --------------------------------------------------------------------------------------------
EMITTER_VERSION = 1
TCM_ABI = 2


def emit_copy(pr: CompiledProblem, fn_id: int, name: str, max_pairs: int = 192, pair: int = 0, wide: int = 0) -> List[Tuple[str, str]]:
    """chunks of one copy function: sequential ``m[d]=m[s]`` (writeCfunctionpergroup.m:753-852)"""
    tab, slot = pr.tab, pr.slot
    c = pr.copies[name]
    pairs = [(d, s) for d, s in zip(c.dest, c.src) if d != s]
    out = []
    for k in range(0, max(len(pairs), 1), max_pairs):
        lines = [f"__device__ __noinline__ void tcc{fn_id}_{k // max_pairs}(double *__restrict__ m, unsigned mask) {{",
                 "TC_TILE_LOOP {", "double *__restrict__ mt = TC_MT;"]
        for d, s in pairs[k:k + max_pairs]:
            ts = tab.types[s]
            if ts == it.I_load:
                src = _lit(tab.params[s][0])
            elif ts == it.I_2:
                src = "2.0"
            else:
                src = f"mt[{int(slot[s]) * TILE * (2 if pair or wide else 1)}]"
            lines.append(f"mt[{int(slot[d]) * TILE * (2 if pair or wide else 1)}]={src};")
        lines += ["}", "}"]
        out.append((f"tcc{fn_id}_{k // max_pairs}", "\n".join(lines) + "\n"))
    return out


def emit_module(gen, max_refs: int = 1536, max_loads: int = 96, part_bytes: int = 400_000, lockstep: int = 1,
                warps_per_sm: int = 8, prefetch: int = 1, merge: int = 24, divcall: int = 0, tiles: int = 1,
                assume_global: int = 0, resident: int = 0, driver: str = "warp", scache: int = 0,
                pair: int = 0, wide: int = 0, fmad: int = 0, roll: int = 0):
    """returns ({file name: source}, stats) of the CUDA module of one generated solver (``fmad`` is a compiler flag
    of build.build_module, an experiment: it does not change the emitted source)"""
    from .csparse import fingerprint
    from .tape import FN
    pr: CompiledProblem = gen.problem
    ext = external_values(pr)
    funcs: List[Tuple[str, str]] = []
    group_calls: List[List[str]] = []
    total = dict(loads=0, stores=0, reg_reads=0, reg_only=0, prefetches=0, smem_loads=0, smem_stores=0, rolled=0)
    per_group = []
    if roll and (scache or pair):
        raise ValueError("roll: rolled runs address the plain scratchbook (no scache, no pair)")
    if scache and ((tiles != 1 and not pair) or resident or driver != "warp"):
        raise ValueError("scache needs the warp driver with one instance per lane (or pair) and no `resident` cap")
    if scache * 8 * TILE * lockstep * (2 if pair else 1) > 227 * 1024:
        raise ValueError("scache: more than 227 KB of shared memory per CTA")
    for g in range(len(pr.groups)):
        fs, st = emit_group(pr, g, ext, max_refs, max_loads, prefetch, merge, assume_global,
                            barrier=int(lockstep > 1), scache=scache, pair=pair, wide=wide, roll=roll)
        funcs += fs
        group_calls.append([n for n, _ in fs])
        per_group.append(st)
        for k in total:
            total[k] += st[k]
    copy_calls: Dict[int, List[str]] = {}
    from .tape import driver_of, fn_items
    if int((pr.defines or {}).get("allowSave", 0)) and driver != "warp":
        raise ValueError("allowSave: the save events are implemented by the warp driver only")
    if driver_of(pr) != "ipmPD_CSsolver" and driver != "warp":
        raise ValueError("the lane-decoupled and work-queue drivers implement ipmPD_CSsolver only")
    for name, fid in fn_items(pr):
        if name in pr.copies:
            fs = emit_copy(pr, fid, name, pair=pair, wide=wide)
            funcs += fs
            copy_calls[fid] = [n for n, _ in fs]
    # part files of roughly equal size (ptxas time is what we balance)
    if wide and (pair or tiles != 1 or lockstep % 2 or driver != "warp"):
        raise ValueError("wide needs the warp driver, one instance per lane and an even lockstep")
    if pair:                       # two interleaved instances per lane: the driver's tiles machinery with the
        if tiles not in (1, 2) or driver != "warp":     # pair layout (TC_PAIR) and 128-bit accesses
            raise ValueError("pair needs the warp driver (and implies tiles=2)")
        tiles = 2
    if tiles not in (1, 2, 4, 8):
        raise ValueError("tiles must be 1, 2, 4 or 8 instances per lane")
    head = ("// generated by tenscalc_b200/emit_cuda.py -- do not edit\n#include <cfloat>\n#include <cmath>\n#include \"tc_math.cuh\"\n"
            f"#define TC_TILES {tiles}\n#define TC_TILE_STRIDE {int(pr.n_slots) * TILE}ull\n"
            + ("// pair: the lane's two instances are interleaved, m[tile pair][slot][lane][2] (tc_image.h, ilv = 1)\n"
               "#define TC_PAIR 1\n"
               '#define TC_TILE_LOOP _Pragma("unroll 1") for (int t = 0; t < 2; t++) if ((mask >> t) & 1u)\n'
               "#define TC_MT (m + t)\n#if 0\n" if pair else "") +
            "#if TC_TILES > 1\n"
            '#define TC_TILE_LOOP _Pragma("unroll 1") for (int t = 0; t < TC_TILES; t++) if ((mask >> t) & 1u)\n'
            "#define TC_MT (m + (size_t)t * TC_TILE_STRIDE)\n"
            "#else\n#define TC_TILE_LOOP if (mask)\n#define TC_MT m\n#endif\n"
            + ("// wide: tiles of 64 instances, m[tile][slot][64]; the lane's column pointer already holds its half\n"
               "#define TC_WIDE 1\n" if wide else "")
            + ("#endif\n" if pair else "") +
            '#define TC_PREFETCH(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))\n'
            "// per-lane software cache in shared memory (emitter option scache, plan_scache): warp w of the CTA owns\n"
            "// TC_SCACHE_SLOTS rows of 32 doubles, slot s of a lane is sc[32*s]\n"
            f"#define TC_SCACHE_SLOTS {scache * (2 if pair else 1)}\n"
            "extern __shared__ double tc_scache[];\n"
            "#define TC_SCACHE (tc_scache + ((threadIdx.x >> 5) * (TC_SCACHE_SLOTS * 32) + (threadIdx.x & 31)"
            + (" * 2" if pair else "") + "))\n"
            "// divcall=1 moves the ~20-instruction IEEE division out of line (-18 % SASS); measured on B200 it\n"
            "// is 5-10 % slower than inline division (the call spills live operands), so inline is the default\n"
            "static __device__ __noinline__ double tc_div(double a, double b) { return a / b; }\n"
            + ("#define TC_DIV(a, b) tc_div((a), (b))\n" if divcall else "#define TC_DIV(a, b) ((a) / (b))\n"))
    files: Dict[str, str] = {}
    cur, size = [head], 0
    for _, src in funcs:
        if size and size + len(src) > part_bytes:
            files[f"part_{len(files):03d}.cu"] = "".join(cur)
            cur, size = [head], 0
        cur.append(src)
        size += len(src)
    files[f"part_{len(files):03d}.cu"] = "".join(cur)
    fp = fingerprint(pr)
    if lockstep not in (1, 2, 3, 4, 6, 8, 12):
        raise ValueError("lockstep must be 1, 2, 3, 4, 6, 8 or 12 warps per CTA (32*lockstep divides the batch stride)")
    if warps_per_sm % lockstep or not 1 <= warps_per_sm <= 32:
        raise ValueError("warps_per_sm must be a multiple of lockstep, at most 32")
    # `resident` > 0 caps the CTAs resident per SM below the register limit (a dynamic shared-memory
    # reservation nobody uses): fewer warps in flight = a larger share of the 126 MB L2 per warp
    smem = 0 if not resident else (227 * 1024 // resident) // 1024 * 1024 - 1024
    if scache:
        smem = scache * 8 * TILE * lockstep * (2 if pair else 1)
    if driver not in ("warp", "lanes", "queue") or (driver != "warp" and (tiles != 1 or lockstep != 1)):
        raise ValueError("driver must be 'warp', 'lanes' or 'queue' (lanes/queue: tiles=1, lockstep=1)")
    m = [head, ("#define TC_DRIVER_LANES 1\n" if driver != "warp" else ""),
         ("#define TC_DRIVER_QUEUE 1\n" if driver == "queue" else ""),
         f"#define TC_COMPILED_GROUPS 1\n#define TC_LOCKSTEP {lockstep}\n#define TC_SMEM_BYTES {smem}\n"
               f"#define TC_KMINBLOCKS {warps_per_sm // lockstep}\n"]
    m += [f"__device__ void {n}(double *__restrict__ m, unsigned mask);\n" for n, _ in funcs]
    m.append("__device__ __noinline__ void tcm_group(int g, double *m, unsigned mask) {\n    switch (g) {\n")
    bar = ""                       # lockstep: the barriers sit in front of every chunk (emit_group)
    for g, calls in enumerate(group_calls):
        m.append(f"    case {g}:" + "".join(f" {n}(m, mask);{bar}" for n in calls) + " break;\n")
    m.append("    }\n}\n")
    m.append("__device__ __noinline__ bool tcm_copy(int fn, double *m, unsigned mask) {\n    switch (fn) {\n")
    for fid, calls in copy_calls.items():
        m.append(f"    case {fid}: " + " ".join(f"{n}(m, mask);" for n in calls) + " return true;\n")
    m.append("    }\n    return false;\n}\n")
    m.append(f'''#include "tc_kernels.cuh"
extern "C" {{
__attribute__((visibility("default"))) int tcm_abi_version(void) {{ return {TCM_ABI}; }}
__attribute__((visibility("default"))) const char *tcm_fingerprint(void) {{ return "{fp}"; }}
__attribute__((visibility("default"))) int tcm_image_bytes(void) {{ return (int)sizeof(tc_device_image); }}
__attribute__((visibility("default"))) int tcm_block(void) {{ return TC_KBLOCK * TC_TILES; }}   // instances per CTA
// work-queue driver: number of resident warps = tiles of the work scratchbook (0: other drivers)
__attribute__((visibility("default"))) int tcm_work_tiles(void) {{
#ifdef TC_DRIVER_QUEUE
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return sms * TC_KMINBLOCKS;
#else
    return 0;
#endif
}}
// force the (lazy) load of the module's device code now rather than inside the first timed solve
__attribute__((visibility("default"))) int tcm_load(void) {{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, tc_solve_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, tc_eval_kernel);
    if (e == cudaSuccess && TC_SMEM_BYTES)
        e = cudaFuncSetAttribute(tc_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e == cudaSuccess && TC_SMEM_BYTES)
        e = cudaFuncSetAttribute(tc_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    // the shared-memory cache must not cost residency: ask for the carve-out that holds TC_KMINBLOCKS CTAs
    if (e == cudaSuccess && TC_SMEM_BYTES) {{
        int pct = (int)((TC_KMINBLOCKS * (TC_SMEM_BYTES + 1024ll) * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 100) pct = 100;
        e = cudaFuncSetAttribute(tc_solve_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_eval_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }}
    return (int)e;
}}
// ipmPD_CSsolver() for every instance of the batch: ONE launch, the Newton loop stays on the device
// (phase_iters > 0: one phase of a phased solve -- at most that many Newton iterations, see tc_driver.cuh)
__attribute__((visibility("default"))) int tcm_solve(const tc_device_image *P, double mu0, int maxIter, double aeU,
                                                     double aeEq, int32_t *status, int32_t *iter, void *stream,
                                                     int phase_first, int phase_iters) {{
    const unsigned grid = P->queue ? (unsigned)P->work_tiles : (unsigned)(P->stride / (TC_KBLOCK * TC_TILES));
    tc_solve_kernel<<<grid, TC_KBLOCK, TC_SMEM_BYTES, (cudaStream_t)stream>>>(
        *P, mu0, maxIter, aeU, aeEq, status, iter, phase_first, phase_iters);
    return (int)cudaGetLastError();
}}
// CTAs of the solve kernel resident per SM as launched (the register budget allows TC_KMINBLOCKS)
__attribute__((visibility("default"))) int tcm_occupancy(void) {{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tc_solve_kernel, TC_KBLOCK, TC_SMEM_BYTES) != cudaSuccess) return -1;
    return n;
}}
// 1: scratchbook layout m[tile pair][slot][lane][2] (emitter option pair), 0: m[tile][slot][lane]
__attribute__((visibility("default"))) int tcm_interleave(void) {{
#if defined(TC_PAIR)
    return 1;
#elif defined(TC_WIDE)
    return 2;
#else
    return 0;
#endif
}}
// 1: the module's driver parks/resumes the solver loop between phase launches (warp driver, one warp per CTA)
__attribute__((visibility("default"))) int tcm_phased(void) {{
#if defined(TC_DRIVER_LANES)
    return 0;
#else
    return 1;
#endif
}}
// lazy evaluation of the parents of get function fn (classname_getOutputs)
__attribute__((visibility("default"))) int tcm_eval(const tc_device_image *P, int fn, void *stream) {{
    tc_eval_kernel<<<(unsigned)(P->stride / (TC_KBLOCK * TC_TILES)), TC_KBLOCK, TC_SMEM_BYTES, (cudaStream_t)stream>>>(*P, fn);
    return (int)cudaGetLastError();
}}
}}
''')
    files["main.cu"] = "".join(m)
    return files, dict(total=total, per_group=per_group, n_functions=len(funcs), fingerprint=fp)
