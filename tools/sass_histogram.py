"""SASS opcode histogram of the shipped liblayoutdetr_sm100.so, per kernel, for the Blackwell-specific instruction families
(tcgen05 MMA = UTCHMMA, TMEM loads = LDTM, TMA = UTMALDG / UTMASTG, mbarriers = SYNCS, tcgen05.commit = UTCBAR).

    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt

Needs only cuobjdump (no GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "layoutdetr_b200", "liblayoutdetr_sm100.so")
FAMILIES = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "UTCBAR", "UTCATOMSWS",
            "SYNCS", "UBLKCP", "HMMA", "MUFU", "REDUX", "LDGSTS", "FENCE", "UTCCP")
INSN = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)")
FUNC = re.compile(r"^\s*Function : (\S+)")
ARCH = re.compile(r"^arch = (\S+)")


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    total = collections.Counter()
    archs = collections.Counter()
    cur = None
    for line in sass.split("\n"):
        m = ARCH.match(line)
        if m:
            archs[m.group(1)] += 1
            continue
        m = FUNC.match(line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = INSN.match(line)
        if m and cur is not None:
            op = m.group(1)
            cur["_all"] += 1
            if op.startswith(FAMILIES):
                cur[op] += 1
                total[op] += 1
    names = demangle(list(per))
    print("# SASS opcode histogram of layoutdetr_b200/liblayoutdetr_sm100.so (cuobjdump -sass), Blackwell-specific families only")
    print("# cubins by arch: %s; %d kernels; %d instructions in total" % (dict(archs), len(per), sum(c["_all"] for c in per.values())))
    print("\n## whole library")
    for op, n in total.most_common():
        print("%8d  %s" % (n, op))
    print("\n## kernels that issue tcgen05 / TMEM / TMA instructions")
    for k, c in per.items():
        if not any(op.startswith(("UTC", "LDTM", "UTMA")) for op in c):
            continue
        short = re.sub(r"^void ", "", names[k])[:110]
        print("\n%s   (%d instructions)" % (short, c["_all"]))
        for op, n in sorted(c.items(), key=lambda t: -t[1]):
            if op != "_all" and not op.startswith(("MUFU", "REDUX", "FENCE")):
                print("%8d  %s" % (n, op))


if __name__ == "__main__":
    sys.exit(main())
