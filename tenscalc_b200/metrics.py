"""Static per-group work tables: the algorithmic flop and byte definitions of the benchmark.

* ``flops_per_group``  operation counts exactly as the reference's C emitter accumulates them
  per emitted instruction (``countFlops[P_nsum..P_cpwise]``,
  ``lib/csparse/instructionsTableUTHash.c:816-832`` and the per-type increments in
  :959-1539, e.g. I_sumprod ``nsum+=s-1, nprod+=(m-1)*s`` :1004-1005); their total is what
  ``examples/flops.m:94-99`` reports as "flops".
* ``boundary_bytes_per_group``  8 bytes for every DISTINCT value a group reads from outside
  itself plus every value it produces that is read outside itself (other groups, gets,
  copies) -- ``groupInputInstructions``/``groupOutputInstructions`` of
  ``lib/@csparse/dependencyGroups.m:236-264``.  This is the traffic that must cross HBM when
  the per-instance scratchbook does not fit on chip but every intra-group temporary does.
* ``vm_bytes_per_group``  8 bytes per scratch operand read and per result write the tape
  interpreter actually issues (what an uncached execution would move).

Multiplied by the per-group execution counts the device collects (``tc_group_counts``) these
give the algorithmic flops / bytes of a run.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from . import itable as it
from .csparse import CompiledProblem


def _flops(t: int, params, n: int) -> int:
    """sum over the flop classes, per instruction type (instructionsTableUTHash.c:959-1539;
    note the integer divisions ``(nOperands-1)/2`` of :1077-1149 are reproduced as written)"""
    if t in (it.I_set, it.I_load, it.I_2):
        return 0
    if t == it.I_sum:
        return n - 1                                   # :962
    if t == it.I_sumprod:
        m, s = int(params[0]), int(params[1])
        return (s - 1) + (m - 1) * s                   # :1004-1005
    if t in (it.I_plus_minus_dot_div, it.I_minus_dot_div):
        return 2 * ((n - 1) // 2) + 1                  # :1077-1079, :1102-1104
    if t in (it.I_plus_minus_dot, it.I_minus_dot):
        return 2 * ((n - 1) // 2)                      # :1125-1126, :1148-1149
    if t in (it.I_plus_sqr, it.I_plus_abs):
        return 2 * n - 1                               # :1198-1199, :1214-1215
    if t in (it.I_min, it.I_max):
        return 2 * (n - 1)                             # :1229-1230
    if t in (it.I_min0, it.I_max0):
        return 2 * n                                   # :1247-1248
    if t == it.I_max_abs:
        return 2 * (n - 1) + n                         # :1297-1299
    if t == it.I_clp:
        return n                                       # :1316
    if t in (it.I_3sqr, it.I_cube, it.I_minus_inv_sqr, it.I_Dsqrt, it.I_DDsqrt, it.I_srelu, it.I_dsrelu):
        return 2
    if t in (it.I_2_inv_cube, it.I_Datan):
        return 3
    if t == it.I_ddsrelu:
        return 4
    if t == it.I_DDatan:
        return 7
    return 1


def group_tables(pr: CompiledProblem) -> Dict[str, np.ndarray]:
    tab = pr.tab
    ng = len(pr.groups)
    flops = np.zeros(ng, dtype=np.int64)
    vm_bytes = np.zeros(ng, dtype=np.int64)
    boundary = np.zeros(ng, dtype=np.int64)
    # values read outside their own group (by another group, a get or a copy)
    external = set()
    for g in pr.gets.values():
        external.update(x for s in g.sources for x in s if x is not None)
    for c in pr.copies.values():
        external.update(c.src)
    for i, gi in pr.group_of.items():
        for o in tab.operands[i]:
            if pr.group_of.get(o, -1) != gi:
                external.add(o)
    const_types = (it.I_load, it.I_2)
    for g, members in enumerate(pr.groups):
        inputs = set()
        mset = set(members)
        for i in members:
            t, p, ops = tab.types[i], tab.params[i], tab.operands[i]
            flops[g] += _flops(t, p, len(ops))
            reads = [o for o in ops if tab.types[o] not in const_types]
            vm_bytes[g] += 8 * (len(reads) + 1)
            inputs.update(o for o in reads if o not in mset)
        outputs = sum(1 for i in members if i in external)
        boundary[g] = 8 * (len(inputs) + outputs)
    return dict(flops=flops, vm_bytes=vm_bytes, boundary_bytes=boundary)


def io_bytes_per_instance(gen) -> Dict[str, int]:
    """bytes that cross the C ABI per instance and solve: parameters + primal initial values
    in, outputs + (status, iter) out"""
    def n(spec):
        return sum(int(np.prod(s)) if s else 1 for _, s in spec)
    return dict(h2d=8 * (n(gen.parameters) + n(gen.variables)), d2h=8 * n(gen.outputs) + 8)
