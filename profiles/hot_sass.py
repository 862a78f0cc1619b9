#!/usr/bin/env python
"""Top stalled SASS instructions per kernel from `ncu --page source --csv
--print-source sass` dumps.  usage: hot_sass.py file.csv [kernel-substring] [top]"""
import csv
import sys


def main(path, want="", top=25):
    rows = list(csv.reader(open(path)))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            hdr = rows[i + 1]
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if len(rows[j]) >= len(hdr) - 2:
                    body.append(dict(zip(hdr, rows[j])))
                j += 1
            i = j
            if want not in name:
                continue
            tot = sum(int(b["# Samples"] or 0) for b in body) or 1
            print(f"== {name[:100]}  ({len(body)} instr, {tot} samples)")
            stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            idx = sorted(range(len(body)), key=lambda k: -int(body[k]["# Samples"] or 0))[:top]
            for k in sorted(idx):
                b = body[k]
                n = int(b["# Samples"] or 0)
                st = sorted(((int(b[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
                print(f"  #{k:5d} {100*n/tot:5.1f}%  {b['Source'][:70]:70s} "
                      + " ".join(f"{c}:{v}" for v, c in st if v))
        else:
            i += 1


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "",
         int(sys.argv[3]) if len(sys.argv) > 3 else 25)
