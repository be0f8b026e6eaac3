"""Opcode histogram (executed warp instructions and stall samples) from `ncu --page source --csv` output."""
import collections
import csv
import sys


def main(path, top=25):
    with open(path) as f:
        rows = list(csv.reader(f))
    hdr = rows[1]
    ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    agg = collections.defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for r in rows[2:]:
        if len(r) <= max(ia, ie, isamp):
            continue
        toks = r[ia].split()
        op = toks[0] if toks and not toks[0].startswith("@") else (toks[1] if len(toks) > 1 else "?")
        op = ".".join(op.split(".")[:2])
        n, s = int(r[ie] or 0), int(r[isamp] or 0)
        agg[op][0] += n
        agg[op][1] += s
        tot_i += n
        tot_s += s
    print(f"total executed warp instructions {tot_i}, stall samples {tot_s}")
    for op, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{op:18s} {n:12d} {100 * n / tot_i:5.1f}%   samples {100 * s / max(tot_s, 1):5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
