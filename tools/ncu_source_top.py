#!/usr/bin/env python
"""Top SASS instructions of an `ncu --page source --csv` export by warp-stall samples, with the
samples aggregated per function, per opcode and per stall reason (profiles/*_source_top.txt).
The export has one section per device function: a "Kernel Name" line, a header line, rows."""
import csv
import re
import sys
from collections import Counter

csv.field_size_limit(1 << 30)


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main(path, top=40):
    hdr, fn = None, "?"
    by_fn, by_op, by_reason = Counter(), Counter(), Counter()
    exec_by_op = Counter()
    best = []                                   # (samples, fn, address, sass, detail)
    window = []
    for r in csv.reader(open(path, errors="replace")):
        if not r:
            continue
        if r[0] == "Kernel Name":
            fn = r[1]
            continue
        if r[0] == "Address":
            if hdr is None:
                print("columns:", r)
            hdr = r
            col = {h: i for i, h in enumerate(hdr)}
            c_src, c_all = col["Source"], col["Warp Stall Sampling (All Samples)"]
            c_exec = col.get("Instructions Executed")
            reasons = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.lower().startswith("stall")]
            continue
        if hdr is None or len(r) <= c_all:
            continue
        s = num(r[c_all])
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[c_src])
        op = m.group(2) if m else "?"
        by_fn[fn] += s
        by_op[op] += s
        if c_exec is not None:
            exec_by_op[op] += num(r[c_exec])
        for i in reasons:
            if i < len(r):
                by_reason[hdr[i]] += num(r[i])
        window.append((fn, r[0], r[c_src][:80], s))
        if len(window) > 6:
            window.pop(0)
        if s > 0:
            detail = " ".join(f"{hdr[i]}={r[i]}" for i in reasons if i < len(r) and num(r[i]) > 0)
            best.append((s, fn, r[0], r[c_src][:80], detail, [w[2] for w in window[:-1]]))
            if len(best) > 4 * top:
                best.sort(key=lambda x: -x[0])
                del best[top:]
    tot = sum(by_fn.values())
    print(f"total samples {tot:.0f}")
    print("by stall reason:", [(k, int(v)) for k, v in by_reason.most_common(20)])
    print("by opcode (samples, executed):", [(k, int(v), int(exec_by_op[k])) for k, v in by_op.most_common(25)])
    print("by function (top 15):", [(k[:30], int(v)) for k, v in by_fn.most_common(15)])
    best.sort(key=lambda x: -x[0])
    for s, f, a, src, detail, before in best[:top]:
        print("-" * 70)
        for b in before[-4:]:
            print("     " + b)
        print(f"=> {src:80s} samples={s:.0f} [{f[:24]} @{a}] {detail[:160]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
