#!/usr/bin/env python3
"""Attribute an `ncu --set full --import-source on` capture of one kernel to the lines of its TOP-LEVEL source file.

    ncu -i prof.ncu-rep --page source --csv --print-source sass > src.csv
    cuobjdump -xelf all lib.so ; nvdisasm -gi file.cubin > all.sass
    tools/ncu_phase_report.py src.csv all.sass <mangled kernel name> <top-level file> [line:label ...]

Every SASS instruction is charged to the OUTERMOST source line of its inlining chain (the last `//## File` note in
front of it), so the work done inside inlined helpers shows up at the kernel-body line that called them.  With
`line:label` pairs, lines are grouped into phases: a label covers lines from its line number up to the next one.
"""
import csv
import re
import sys
from collections import defaultdict


def sass_line_map(path, kernel, top):
    """offset (int) -> outermost line of `top` (or (file, line))."""
    m, on, last = {}, False, None
    note = re.compile(r'//## File "([^"]+)", line (\d+)')
    ins = re.compile(r'^\s+/\*([0-9a-f]{4,6})\*/')
    with open(path) as f:
        for ln in f:
            if ln.startswith(".text."):
                on = ln.strip().rstrip(":") == ".text." + kernel
                last = None
                continue
            if not on:
                continue
            t = note.search(ln)
            if t:
                last = (t.group(1), int(t.group(2)))  # the FIRST group of the last note = outermost frame
                continue
            t = ins.match(ln)
            if t and last is not None:
                m[int(t.group(1), 16)] = last
    return m


def main():
    src_csv, sass, kernel, top = sys.argv[1:5]
    labels = sorted((int(a.split(":")[0]), a.split(":", 1)[1]) for a in sys.argv[5:])
    lm = sass_line_map(sass, kernel, top)
    rows = list(csv.reader(open(src_csv)))
    # the file may hold several captured launches: take the first block
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or r[0] in ("Address", "Kernel Name"):
            break
        body.append(r)
    base = int(body[0][0], 16)
    keys = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "stall_barrier", "stall_short_sb",
            "stall_long_sb", "stall_wait", "stall_math", "stall_mio", "stall_not_selected", "stall_selected", "stall_dispatch"]
    agg = defaultdict(lambda: defaultdict(float))
    ops = defaultdict(lambda: defaultdict(float))
    for r in body:
        off = int(r[0], 16) - base
        f, line = lm.get(off, ("?", 0))
        key = line if f.endswith(top) else -1
        if labels and key >= 0:
            lab = "pre"
            for ln, name in labels:
                if key >= ln:
                    lab = name
            key = lab
        for k in keys:
            try:
                agg[key][k] += float(r[col[k]])
            except (ValueError, KeyError):
                pass
        op = r[1].split()[0] if r[1].split() else "?"
        if op.startswith("@"):
            op = r[1].split()[1]
        op = op.split(".")[0]
        try:
            ops[key][op] += float(r[col["Instructions Executed"]])
        except ValueError:
            pass
    tot = defaultdict(float)
    for k in agg:
        for kk, v in agg[k].items():
            tot[kk] += v
    print("%-22s %8s %6s %10s %10s | %7s %7s %7s %7s %7s %7s | top opcodes (warp-instr)" %
          ("line/phase", "samples", "%", "warp-inst", "smem wf", "barrier", "shortsb", "longsb", "wait", "math", "mio"))
    order = [n for _, n in labels] if labels else sorted(agg, key=lambda x: (isinstance(x, str), x))
    if labels:
        order = ["pre"] + order + [k for k in agg if k not in order and k != "pre"]
    for k in order:
        if k not in agg:
            continue
        a = agg[k]
        top_ops = sorted(ops[k].items(), key=lambda kv: -kv[1])[:6]
        print("%-22s %8d %6.1f %10d %10d | %7d %7d %7d %7d %7d %7d | %s" % (
            str(k), a["# Samples"], 100.0 * a["# Samples"] / max(1.0, tot["# Samples"]), a["Instructions Executed"], a["L1 Wavefronts Shared"],
            a["stall_barrier"], a["stall_short_sb"], a["stall_long_sb"], a["stall_wait"], a["stall_math"], a["stall_mio"],
            " ".join("%s:%d" % (o, n) for o, n in top_ops)))
    print("%-22s %8d %6s %10d %10d" % ("total", tot["# Samples"], "", tot["Instructions Executed"], tot["L1 Wavefronts Shared"]))


if __name__ == "__main__":
    main()
