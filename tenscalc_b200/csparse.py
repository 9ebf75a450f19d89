"""Declarations (set/get/copy), dependency groups and memory map of one problem.

Host-side, offline.  Counterpart of the reference's ``lib/@csparse`` object as far
as the IPM hot path needs it:

* ``declareSet/declareGet/declareCopy`` (``csparse.m:296-650``),
* dependency groups (``dependencyGroups.m:36-224``): an instruction's signature is
  the set of gets/copies that need it and of sets/copies it depends on;
  instructions with equal signature form one group, groups are topologically
  ordered and re-evaluated lazily through one status flag per group
  (``writeCfunctionpergroup.m:44,566-571,653-657,843-846``),
* the scratchbook memory map (``dependencyGroups.m:275-422``).

The result (:class:`CompiledProblem`) is what both back ends consume: the device
tape writer (``tape.py``) and the C emitters of the CPU oracle (``oracle/``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import itable as it
from .itable import InstructionTable, Id


@dataclass
class SetDecl:
    name: str
    dest: List[int]                 # I_set instructions, column-major
    shape: Tuple[int, ...]
    kind: str = "solver"            # "parameter" | "variable" | "solver"
    children: List[int] = field(default_factory=list)   # group ids


@dataclass
class GetDecl:
    name: str
    sources: List[List[Id]]         # one list per output, column-major, None = structural 0
    shapes: List[Tuple[int, ...]]
    out_names: List[str] = field(default_factory=list)
    parents: List[int] = field(default_factory=list)


@dataclass
class CopyDecl:
    name: str
    dest: List[int]
    src: List[int]
    parents: List[int] = field(default_factory=list)
    children: List[int] = field(default_factory=list)


@dataclass
class CompiledProblem:
    """Everything a back end needs; instruction ids are table ids (0-based)."""
    tab: InstructionTable
    sets: Dict[str, SetDecl]
    gets: Dict[str, GetDecl]
    copies: Dict[str, CopyDecl]
    groups: List[List[int]]                 # groups[g] = instruction ids in execution order
    group_of: Dict[int, int]
    slot: np.ndarray                        # slot[id] = scratchbook location or -1
    n_slots: int
    defines: Dict[str, object] = field(default_factory=dict)
    flops_per_group: List[int] = field(default_factory=list)


def fingerprint(pr: CompiledProblem) -> str:
    """hash of the exact instruction list, memory map, groups and #define table: two problems
    with the same fingerprint compile to the same straight-line code on every back end"""
    import hashlib
    h = hashlib.sha1()
    tab = pr.tab
    for members in pr.groups:
        h.update(np.asarray(members, dtype=np.int64).tobytes())
        for i in members:
            h.update(repr((tab.types[i], tab.params[i], tab.operands[i])).encode())
    h.update(np.asarray(pr.slot, dtype=np.int64).tobytes())
    h.update(repr(sorted(pr.defines.items())).encode())
    for d in (pr.sets, pr.gets, pr.copies):
        for k, v in d.items():
            h.update(repr((k, getattr(v, "dest", None), getattr(v, "src", None), getattr(v, "sources", None))).encode())
    return h.hexdigest()


class CSparse:
    def __init__(self, tab: InstructionTable):
        self.tab = tab
        self.sets: Dict[str, SetDecl] = {}
        self.gets: Dict[str, GetDecl] = {}
        self.copies: Dict[str, CopyDecl] = {}

    # -- declarations -----------------------------------------------------------------
    def declare_set(self, name: str, dest: Sequence[int], shape=(), kind: str = "solver") -> None:
        tab = self.tab
        for d in dest:
            if d is None or tab.types[d] != it.I_set:
                raise ValueError(f"declareSet({name}): destination must be a full variable")
        self.sets[name] = SetDecl(name, list(dest), tuple(shape), kind)

    def declare_get(self, name: str, sources: Sequence[Sequence[Id]], shapes=None, out_names=None) -> None:
        srcs = [list(s) for s in sources]
        if shapes is None:
            shapes = [(len(s),) if len(s) != 1 else () for s in srcs]
        self.gets[name] = GetDecl(name, srcs, [tuple(s) for s in shapes], list(out_names or []))

    def declare_copy(self, name: str, dest: Sequence[int], src: Sequence[Id]) -> None:
        tab = self.tab
        if len(dest) != len(src):
            raise ValueError(f"declareCopy({name}): size mismatch")
        for d, s in zip(dest, src):
            if tab.types[d] != it.I_set:
                raise ValueError(f"declareCopy({name}): destination must be a variable")
            if s is None:
                # the reference raises here too: source and destination sparsity differ
                # (writeCfunctionpergroup.m:808-826)
                raise ValueError(f"declareCopy({name}): source has structural zeros, use full()")
        self.copies[name] = CopyDecl(name, list(dest), list(src))  # type: ignore[arg-type]

    # -- compilation ------------------------------------------------------------------
    def compile(self, defines: Optional[Dict[str, object]] = None) -> CompiledProblem:
        tab = self.tab
        n = len(tab)
        types = np.asarray(tab.types, dtype=np.int32)

        # one signature column per get / copy-source (ancestors) ...
        get_roots: List[Tuple[str, List[int]]] = []
        for g in self.gets.values():
            get_roots.append((g.name, [x for s in g.sources for x in s if x is not None]))
        for c in self.copies.values():
            get_roots.append((c.name, list(c.src)))
        # ... and per set / copy-destination (descendants).  All setP_* share one column
        # and all setV_* share another: finer invalidation buys nothing for a batched
        # solve (every parameter is uploaded before `solve`) and keeps the number of
        # groups small enough for a one-word status mask per instance.
        set_cols: Dict[str, List[int]] = {}
        for s in self.sets.values():
            key = {"parameter": "setP_*", "variable": "setV_*"}.get(s.kind, s.name)
            set_cols.setdefault(key, []).extend(s.dest)
        for c in self.copies.values():
            set_cols[c.name + ">dst"] = list(c.dest)

        # signatures as bit masks, one forward and one backward sweep over the table
        # (instruction ids are a topological order: operands precede their users)
        ops = tab.operands
        need = [0] * n                       # bit k: some root of get/copy k needs it
        for k, (_, roots) in enumerate(get_roots):
            for r in roots:
                need[r] |= 1 << k
        for i in range(n - 1, -1, -1):
            m = need[i]
            if m:
                for o in ops[i]:
                    need[o] |= m
        desc = [0] * n                       # bit k: depends on set/copy-destination k
        for k, dests in enumerate(set_cols.values()):
            for d in dests:
                desc[d] |= 1 << k
        for i in range(n):
            if need[i]:
                m = desc[i]
                for o in ops[i]:
                    m |= desc[o]
                desc[i] = m
        live = np.array([m != 0 for m in need], dtype=bool)

        ids = [i for i in range(n) if need[i] and types[i] != it.I_set]
        sig_index: Dict[Tuple[int, int], int] = {}
        members: List[List[int]] = []
        gid_of: Dict[int, int] = {}
        for i in ids:
            key = (need[i], desc[i])
            g = sig_index.get(key)
            if g is None:
                g = sig_index[key] = len(members)
                members.append([])
            members[g].append(i)
            gid_of[i] = g
        ngroups = len(members)

        # topological order of groups
        succ = [set() for _ in range(ngroups)]
        indeg = [0] * ngroups
        for i in ids:
            gi = gid_of[i]
            for o in ops[i]:
                go = gid_of.get(o)
                if go is not None and go != gi and gi not in succ[go]:
                    succ[go].add(gi)
                    indeg[gi] += 1
        ready = sorted(g for g in range(ngroups) if indeg[g] == 0)
        order: List[int] = []
        while ready:
            g = ready.pop(0)
            order.append(g)
            for h in sorted(succ[g]):
                indeg[h] -= 1
                if indeg[h] == 0:
                    ready.append(h)
            ready.sort()
        if len(order) != ngroups:
            raise RuntimeError("dependency groups have a cycle")
        rank = {g: r for r, g in enumerate(order)}
        groups = [members[g] for g in order]
        group_of = {i: rank[g] for i, g in gid_of.items()}

        # parents / children of every declaration, in group (= execution) order
        group_need = [0] * ngroups
        group_desc = [0] * ngroups
        for (nd, ds), g in sig_index.items():
            group_need[rank[g]] = nd
            group_desc[rank[g]] = ds
        root_bit = {name: k for k, (name, _) in enumerate(get_roots)}
        set_bit = {name: k for k, name in enumerate(set_cols)}

        def parents_of(name: str) -> List[int]:
            b = 1 << root_bit[name]
            return [g for g in range(ngroups) if group_need[g] & b]

        def children_of(colname: str) -> List[int]:
            b = 1 << set_bit[colname]
            return [g for g in range(ngroups) if group_desc[g] & b]

        for g in self.gets.values():
            g.parents = parents_of(g.name)
        for c in self.copies.values():
            c.parents = parents_of(c.name)
            c.children = children_of(c.name + ">dst")
        for s in self.sets.values():
            key = {"parameter": "setP_*", "variable": "setV_*"}.get(s.kind, s.name)
            s.children = children_of(key)

        slot, n_slots = self._memory_map(live, groups, group_of)
        return CompiledProblem(tab, dict(self.sets), dict(self.gets), dict(self.copies),
                               groups, group_of, slot, n_slots, dict(defines or {}))

    # ---------------------------------------------------------------------------------
    def _memory_map(self, live: np.ndarray, groups: List[List[int]], group_of: Dict[int, int]):
        """Scratchbook slots.  Variables first (each declared variable contiguous, so
        that sets are block copies), then group results.  A value that is only read
        inside its own group and is not a get/copy source is a local temporary and its
        slot is recycled after its last use (dependencyGroups.m:310-396)."""
        tab = self.tab
        n = len(tab)
        slot = -np.ones(n, dtype=np.int64)
        nxt = 0
        for s in self.sets.values():
            for d in s.dest:
                if slot[d] < 0:
                    slot[d] = nxt
                    nxt += 1
        for i in np.nonzero(live)[0]:
            if tab.types[i] == it.I_set and slot[i] < 0:
                slot[i] = nxt
                nxt += 1
        pinned = set()
        for g in self.gets.values():
            pinned.update(x for s in g.sources for x in s if x is not None)
        for c in self.copies.values():
            pinned.update(c.src)
        # last use inside the group / used outside the group
        external = set(pinned)
        for i, gi in group_of.items():
            for o in tab.operands[i]:
                if group_of.get(o, -1) != gi:
                    external.add(o)
        free: List[int] = []          # temporary slots, shared by all groups: a group runs
        for members in groups:        # to completion before any other group starts
            last_use: Dict[int, int] = {}
            for pos, i in enumerate(members):
                for o in tab.operands[i]:
                    last_use[o] = pos
            release_at: Dict[int, List[int]] = {}
            for pos, i in enumerate(members):
                if i in external:
                    slot[i] = nxt
                    nxt += 1
                else:
                    if free:
                        slot[i] = free.pop()
                    else:
                        slot[i] = nxt
                        nxt += 1
                    release_at.setdefault(last_use.get(i, pos), []).append(i)
                # temporaries whose last reader is this instruction become available to
                # the NEXT instruction (a result never aliases one of its own operands)
                for r in release_at.pop(pos, []):
                    free.append(int(slot[r]))
        return slot, int(nxt)
