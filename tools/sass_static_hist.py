"""Static SASS opcode histogram of the kernels in the built library whose (demangled) names match a pattern:

  python tools/sass_static_hist.py <pattern> [max_kernels]

Reads `cuobjdump -sass` of multi-modal-emotion-recognition_b200/libmmer_sm100.so (no GPU needed).  What to look for
(B200_PROFILING.md): UTCHMMA / UTCBAR = tcgen05.mma / commit, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UBLKCP = TMA,
HMMA = legacy mma.sync, LDGMC / multimem = NVSwitch multicast loads."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multi-modal-emotion-recognition_b200", "libmmer_sm100.so")


def main(pattern, limit):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    name, kernels = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
    shown = 0
    for k, ops in kernels.items():
        d = demangle(k)
        if not re.search(pattern, d):
            continue
        tot = sum(ops.values())
        print(f"== {d[:200]}\n   {tot} SASS instructions")
        keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "HMMA", "LDSM", "MOVM", "LDGMC",
                "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "BAR", "ATOMG", "RED", "REDG", "LDG", "STG", "LDS", "STS", "MUFU", "SHFL"]
        fam = collections.Counter()
        for op, n in ops.items():
            fam[op.split(".")[0]] += n
        print("   " + "  ".join(f"{k}:{fam[k]}" for k in keys if fam.get(k)))
        print("   top: " + "  ".join(f"{op}:{n}" for op, n in fam.most_common(14)))
        shown += 1
        if shown >= limit:
            break


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 6)
