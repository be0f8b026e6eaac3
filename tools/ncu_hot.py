"""Top stall sites of a kernel from an .ncu-rep source page (SASS level, with the owning source line when -lineinfo).

  python tools/ncu_hot.py report.ncu-rep [N]
"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name",')
    for blk in blocks[1:]:
        lines = blk.split("\n")
        print("==", lines[0][:120])
        rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        hdr = rows[0]
        si = hdr.index("# Samples") if "# Samples" in hdr else None
        src = hdr.index("Source")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        data = []
        total = 0
        for r in rows[1:]:
            if len(r) <= si:
                continue
            try:
                s = int(r[si])
            except ValueError:
                continue
            total += s
            data.append((s, r))
        data_sorted = sorted(data, key=lambda x: -x[0])[:n]
        print(f"total samples {total}")
        for s, r in data_sorted:
            reasons = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
            rs = " ".join(f"{nm}:{v}" for v, nm in reasons if v)
            print(f"{s:7d} {100.0 * s / max(total, 1):5.1f}%  {r[src][:90]:90s} {rs}")


if __name__ == "__main__":
    main()
