#!/usr/bin/env python
"""Instruction mix of a kernel's hottest loop, from `cuobjdump -sass`.

    python tools/sass_inner_loop.py cardelino_b200/build/cells_sell.o Li16EtLb1ELi896ELi1ELb0ELb1 [--list]

Finds the function whose mangled name contains the given substring, takes the backward branch whose body has the
highest share of DADD instructions (the accumulate loop of k_cells_sell), and prints the instruction count per
mnemonic and per covered entry (one DMUL per entry).  --list also prints the loop body.  Used to produce
profiles/r2_sass_k_cells_sell_inner_loop.txt; needs no GPU."""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    listing = "--list" in sys.argv
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout.split("\n")
    start = next(i for i, l in enumerate(sass) if "Function" in l and pat in l)
    end = next((i for i in range(start + 1, len(sass)) if "Function" in sass[i]), len(sass))
    ins = []
    for l in sass[start:end]:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    at = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:[!\w]+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in at:
            body = ins[at[int(m.group(1), 16)]:i + 1]
            nd = sum("DADD" in x for _, x in body)
            if nd >= 50 and (best is None or nd / len(body) > best[0]):
                best = (nd / len(body), body)
    if best is None:
        raise SystemExit("no DADD loop found")
    body = best[1]
    entries = max(1, sum("DMUL" in x for _, x in body))
    mix = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for _, t in body)
    print("%s" % sass[start].strip())
    print("loop %05x..%05x: %d instructions, %d covered entries per lane, %.2f instructions per entry" %
          (body[0][0], body[-1][0], len(body), entries, len(body) / entries))
    for k, v in mix.most_common():
        print("  %-8s %4d  %5.2f per entry" % (k, v, v / entries))
    if listing:
        print()
        for a, t in body:
            print("%05x  %s" % (a, t))


if __name__ == "__main__":
    main()
