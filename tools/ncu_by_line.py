#!/usr/bin/env python3
"""Aggregate an ncu SASS source page by CUDA source line.

    ncu -i prof.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all libptb200.so ; nvdisasm -g -c integrator.sm_100a.cubin > dis.txt
    python tools/ncu_by_line.py sass.csv dis.txt '<mangled kernel name>' [top N]

Joins on the instruction offset inside the function; prints, per source line: warp instructions executed,
share of the kernel's issue slots, average active threads, and stall samples.
"""
import csv
import re
import sys
from collections import defaultdict


def parse_disasm(path, func):
    lines = open(path, errors="replace").read().split("\n")
    out = {}
    cur_line, in_func = None, False
    for ln in lines:
        if ln.startswith("\t.section") or ln.startswith(".section"):
            in_func = (".text." + func) in ln
            continue
        if not in_func:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
    return out


def main():
    sass_csv, dis, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dmap = parse_disasm(dis, func)
    rows = list(csv.reader(open(sass_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    data = rows[hdr_i + 1:]
    base = int(data[0][0], 16)
    agg = defaultdict(lambda: [0, 0, 0, 0])   # inst, thread inst, samples, not-issued samples
    tot_i = tot_t = tot_s = 0
    miss = 0
    for r in data:
        off = int(r[0], 16) - base
        inst = int(r[col["Instructions Executed"]] or 0)
        tin = int(r[col["Thread Instructions Executed"]] or 0)
        smp = int(r[col["# Samples"]] or 0)
        key = dmap.get(off, (None, ""))[0]
        if key is None:
            miss += 1
            key = ("?", 0, False)
        a = agg[(key[0], key[1])]
        a[0] += inst; a[1] += tin; a[2] += smp
        tot_i += inst; tot_t += tin; tot_s += smp
    print(f"total warp-inst {tot_i:.4g}, thread-inst {tot_t:.4g}, avg active {tot_t / max(tot_i, 1):.2f}, samples {tot_s}, unmapped sass {miss}")
    print(f"{'file:line':28s} {'warp-inst':>12s} {'%issue':>7s} {'act.thr':>7s} {'%samples':>8s}")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f + ':' + str(l):28s} {a[0]:12d} {100 * a[0] / tot_i:7.2f} {a[1] / max(a[0], 1):7.2f} {100 * a[2] / max(tot_s, 1):8.2f}")


if __name__ == "__main__":
    main()
