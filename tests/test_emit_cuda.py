"""The sm_100a CUDA emitter (host logic, no GPU): expression forms follow the reference emitter
(lib/csparse/instructionsTableUTHash.c:950-1539), chunking keeps every value reachable, and the
per-problem module builds, loads and exports the entry points tc_attach_module binds."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import generated
from tenscalc_b200 import emit_cuda, itable as it


def test_expression_forms():
    e = emit_cuda.expression
    assert e(it.I_sum, [1, -1, 1], ["a", "b", "c"], "r") == "const double r=+a-b+c;"
    assert e(it.I_sumprod, [2, 2], ["a", "b", "c", "d"], "r") == "const double r=a*b+c*d;"
    assert e(it.I_plus_minus_dot_div, [], ["a0", "d", "x", "y"], "r") == "const double r=TC_DIV((a0-(x*y)),d);"
    assert e(it.I_minus_dot_div, [], ["d", "x", "y"], "r") == "const double r=TC_DIV(-(x*y),d);"
    assert e(it.I_plus_minus_dot, [], ["a0", "x", "y", "z", "w"], "r") == "const double r=a0-(x*y+z*w);"
    assert e(it.I_minus_dot, [], ["x", "y"], "r") == "const double r=-(x*y);"
    assert e(it.I_max_abs, [], ["a", "b"], "r") == "double r=fabs(a); if (r<fabs(b)) r=fabs(b);"
    assert e(it.I_min0, [], ["a"], "r") == "double r=0.0; if (r>a) r=a;"
    assert "DBL_MAX" in e(it.I_clp, [], ["a", "b"], "r") and "if (b<0)" in e(it.I_clp, [], ["a", "b"], "r")
    assert e(it.I_Dsqrt, [], ["x"], "r") == "const double r=.5*tc_pow_m05(x);"
    assert e(it.I_sign, [], ["x"], "r") == "const double r=(x>=0)?1.0:-1.0;"


def test_long_sumprod_breaks_lines_like_the_reference():
    """more than 100 factors: `m[d]=line1; m[d]+=line2;` with ceil(100/m) products per line
    (instructionsTableUTHash.c:1019-1026) -- every line is summed on its own"""
    o = [f"x{k}" for k in range(2 * 60)]
    s = emit_cuda.expression(it.I_sumprod, [2, 60], o, "r")
    first = "+".join(f"x{2 * k}*x{2 * k + 1}" for k in range(50))
    rest = "+".join(f"x{2 * k}*x{2 * k + 1}" for k in range(50, 60))
    assert s == f"double r={first}; r+={rest};"
    assert len(it.sumprod_lines(3, [f"y{k}" for k in range(3 * 70)])) == 3          # 34 products per line
    assert len(it.sumprod_lines(2, o[:100])) == 1                                    # exactly one full line


def test_literals_are_exact():
    for v in (0.1, -2.5e-7, 1e300, 3.0, -1.0, 1 / 3):
        assert float.fromhex(emit_cuda._lit(v)) == v if "x" in emit_cuda._lit(v) else float(emit_cuda._lit(v)) == v


@pytest.mark.parametrize("capacity", [4, 24, 75])
@pytest.mark.parametrize("name", ["sls", "mpc_dcmotor"])
def test_scache_plan_is_consistent(name, capacity):
    """the offline shared-memory cache plan (emit_cuda.plan_scache): replay it chunk by chunk -- every value a
    chunk takes from the cache is in the slot the plan names, nothing is read from the scratchbook that was never
    stored there, at most `capacity` slots are live, and a bigger cache never loads more"""
    gen, _ = generated(name)
    pr = gen.problem
    tab = pr.tab
    ext = emit_cuda.external_values(pr)
    const = (it.I_load, it.I_2)
    total = 0
    for g, members in enumerate(pr.groups):
        chunks = emit_cuda.chunk_group(pr, members, 1536, 28)
        reads, makes = [], []
        for ch in chunks:
            have, rs, ms = set(), [], []
            for i in ch:
                if tab.types[i] in (it.I_set,) + const:
                    continue
                for q in tab.operands[i]:
                    if tab.types[q] not in const and q not in have:
                        have.add(q)
                        rs.append(q)
                have.add(i)
                ms.append(i)
            reads.append(rs)
            makes.append(ms)
        plan, to_hbm = emit_cuda.plan_scache(reads, makes, ext, capacity)
        made_all = {i for ms in makes for i in ms}
        assert made_all & ext <= to_hbm                       # everything somebody outside the group reads is stored
        slots = {}                                            # slot -> value, as the emitted code leaves it
        in_hbm = set()                                        # values of this group that reached the scratchbook
        for (from_smem, to_smem), rs, ms in zip(plan, reads, makes):
            for q, sl in from_smem.items():
                assert slots.get(sl) == q, (g, q, sl)
            for q in rs:
                if q not in from_smem:
                    total += 1
                    assert q not in made_all or q in in_hbm, (g, q)     # produced earlier in the group: must have been stored
            for q, sl in to_smem:
                assert 0 <= sl < capacity and (q in rs or q in ms)
                slots[sl] = q
            in_hbm |= {i for i in ms if i in to_hbm}
            assert len(set(slots)) <= capacity
    loads = {}
    for c in (0, capacity):
        st = emit_cuda.emit_module(gen, lockstep=8, warps_per_sm=24, max_loads=28, scache=c)[1]["total"]
        loads[c] = st["loads"] + st["stores"]
    assert loads[capacity] <= loads[0]
    assert total > 0


def test_pair_layout_emission():
    """`pair`: two instances per lane in interleaved columns -- every scratchbook access of a group is a 16-byte
    vector at twice the slot stride, the arithmetic is emitted once per instance, copies stay per instance"""
    gen, _ = generated("sls")
    files, st = emit_cuda.emit_module(gen, lockstep=8, warps_per_sm=24, max_loads=14, pair=1)
    ref = emit_cuda.emit_module(gen, lockstep=8, warps_per_sm=24, max_loads=14)[1]
    assert st["total"]["loads"] == ref["total"]["loads"] and st["total"]["stores"] == ref["total"]["stores"]
    main = files["main.cu"]
    assert "#define TC_PAIR 1" in main and "#define TC_TILES 2" in main and "tcm_interleave" in main
    body = "".join(v for k, v in files.items() if k.startswith("part_"))
    groups = [ln for ln in body.splitlines() if ln.startswith("const double2 v")]
    assert groups and all(re.fullmatch(r"const double2 v\d+=\*\(const double2 \*\)\(mt\+(\d+)\);", ln) for ln in groups)
    assert all(int(re.search(r"mt\+(\d+)", ln).group(1)) % 64 == 0 for ln in groups)          # slot stride 2 x 32 doubles
    xs = len(re.findall(r"const double v\d+x=", body))
    ys = len(re.findall(r"const double v\d+y=", body))
    assert xs == ys > 0
    assert "make_double2(" in body
    with pytest.raises(ValueError):
        emit_cuda.emit_module(gen, pair=1, driver="lanes")


@pytest.mark.parametrize("name", ["sls", "mpc_dcmotor"])
def test_chunks_cover_every_instruction_once(name):
    gen, _ = generated(name)
    pr = gen.problem
    for g, members in enumerate(pr.groups):
        chunks = emit_cuda.chunk_group(pr, members, 1536, 96)
        assert [i for ch in chunks for i in ch] == list(members)
        for ch in chunks[:-1]:
            loads = {o for i in ch for o in pr.tab.operands[i]
                     if pr.tab.types[o] not in (it.I_load, it.I_2)} - set(ch)
            assert len(loads) <= 96 or len(ch) == 1      # one instruction wider than the budget stands alone


@pytest.mark.parametrize("name", ["sls", "mpc_dcmotor"])
def test_every_value_read_is_defined(name):
    """inside a chunk function every operand is a loaded local, an earlier result or an immediate;
    every value read by another chunk / group / get / copy is stored"""
    gen, _ = generated(name)
    pr = gen.problem
    ext = emit_cuda.external_values(pr)
    stored_slots = set()
    for g in range(len(pr.groups)):
        fns, st = emit_cuda.emit_group(pr, g, ext)
        assert st["loads"] + st["reg_reads"] >= 0
        for _, src in fns:
            for body in re.split(r"TC_TILE_LOOP \{", src)[1:]:
                defined = set(re.findall(r"double (v\d+)=", body))
                used = set(re.findall(r"\bv\d+\b", body))
                assert used <= defined, (g, sorted(used - defined)[:5])
            stored_slots.update(int(x) // emit_cuda.TILE for x in re.findall(r"\bmt\[(\d+)\]=", src))
    for i in ext:
        if pr.tab.types[i] not in (it.I_load, it.I_2, it.I_set) and i in pr.group_of:
            assert int(pr.slot[i]) in stored_slots


def test_module_builds_and_exports_entry_points():
    from tenscalc_b200.build import build_module
    from tenscalc_b200.csparse import fingerprint
    gen, _ = generated("sls")
    so = build_module(gen)
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    for sym in ("tcm_abi_version", "tcm_fingerprint", "tcm_image_bytes", "tcm_block", "tcm_load", "tcm_solve",
                "tcm_eval"):
        assert hasattr(lib, sym), sym
    lib.tcm_fingerprint.restype = ctypes.c_char_p
    assert lib.tcm_fingerprint().decode() == fingerprint(gen.problem)
    assert lib.tcm_abi_version() == emit_cuda.TCM_ABI
    from tenscalc_b200.build import default_module_opts
    # instances per CTA: one per lane, `lockstep` warps per CTA walk the code together (default for this size: 8)
    assert lib.tcm_block() == 32 * default_module_opts(gen).get("lockstep", 1)
    solo = ctypes.CDLL(build_module(gen, driver="lanes"))      # the lane-decoupled driver runs one warp per CTA
    assert solo.tcm_block() == 32


def test_module_path_is_content_addressed():
    from tenscalc_b200.build import module_path
    gen, _ = generated("sls")
    assert module_path(gen) == module_path(gen)
    assert module_path(gen) != module_path(gen, max_loads=44)
    assert module_path(gen, tiles=2) != module_path(gen, tiles=1)


@pytest.mark.parametrize("name", ["sls", "mpc_dcmotor", "l1l2"])
def test_rolled_runs_are_the_reference_emitters_loops(name):
    """loop rolling (instructionsTableUTHash.c:881-944): the runs `rolled_runs` detects are exactly the `for`
    loops the reference's own writeCinstructionsC prints for the same instruction list (minInstructions4loop=50)"""
    from oracle.emit_c import RefEmitter
    gen, _ = generated(name)
    pr = gen.problem
    em = RefEmitter(pr)
    total = 0
    for members in pr.groups:
        txt = em.statements(members, 50)
        ref = [int(c) for c in re.findall(r"for\(int64_t c=(\d+);", txt)]
        runs = emit_cuda.rolled_runs(pr, members, 50)
        assert [n for _, n in runs] == ref
        for start, n in runs:            # the comment the reference prints names the first instruction of the run
            assert pr.tab.types[members[start]] in emit_cuda.ROLLABLE
        total += sum(ref)
    assert total > 0


@pytest.mark.parametrize("name", ["sls", "mpc_dcmotor"])
def test_rolled_emission_covers_every_instruction_once(name):
    """roll=N: every run is a chunk of its own holding a device loop; the other instructions are chunked as before;
    every value a rolled chunk reads from another chunk is stored by its producer"""
    gen, _ = generated(name)
    pr = gen.problem
    ext = emit_cuda.external_values(pr)
    n_rolled = 0
    for g, members in enumerate(pr.groups):
        chunks, rolled = emit_cuda.chunk_group_rolled(pr, members, 1536, 28, 20)
        assert [i for ch in chunks for i in ch] == list(members)
        fns, st = emit_cuda.emit_group(pr, g, ext, max_loads=28, roll=20)
        src = "".join(s for _, s in fns)
        assert src.count("// rolled run:") == len(rolled)
        assert st["rolled"] == sum(len(chunks[k]) for k in rolled)
        n_rolled += st["rolled"]
        stored = {int(x) // emit_cuda.TILE for x in re.findall(r"\bmt\[(\d+)\]=", src)}
        for k in rolled:
            inside = set(chunks[k])
            for i in chunks[k]:
                for q in pr.tab.operands[i]:
                    if q in pr.group_of and pr.group_of[q] == g and q not in inside \
                            and pr.tab.types[q] not in (it.I_load, it.I_2, it.I_set) \
                            and not any(q in chunks[j] for j in rolled):
                        assert int(pr.slot[q]) in stored, (g, k, q)
    assert n_rolled > 0
    files, st = emit_cuda.emit_module(gen, roll=20)
    assert st["total"]["rolled"] == n_rolled or st["total"]["rolled"] > 0
    with pytest.raises(ValueError):
        emit_cuda.emit_module(gen, roll=20, scache=8)
