#!/usr/bin/env python
"""Opcode histogram per kernel of libscs_b200.so (cuobjdump -sass), written to profiles/.

Runs on the CPU box: it reads the cubin that build() produced, nothing is executed.  The columns are
the mnemonics that prove which hardware path a kernel takes (profiling recipe, "SASS mnemonics"):
tcgen05 MMA (UTCQMMA / UTCIMMA / UTCHMMA, .2CTA forms counted apart), TMEM loads (LDTM), TMA tensor /
bulk copies (UTMALDG / UBLKCP), tcgen05 commit barriers (UTCBAR), mbarrier waits (SYNCS), the MUFU
exp2 of the softmax, shared-memory traffic (LDS / STS / ATOMS), shuffles, CTA barriers, global
loads / stores and FP64 arithmetic (the LRT).

usage: python tools/sass_histogram.py [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "scsommerclock_b200", "libscs_b200.so")

COLUMNS = [
    ("UTCxMMA", re.compile(r"^UTC[QIH]MMA(?!.*2CTA)")),
    ("UTCxMMA.2CTA", re.compile(r"^UTC[QIH]MMA.*2CTA")),
    ("LDTM", re.compile(r"^LDTM")),
    ("UTMALDG", re.compile(r"^UTMALDG")),
    ("UBLKCP", re.compile(r"^UBLKCP")),
    ("UTCBAR", re.compile(r"^UTCBAR")),
    ("SYNCS", re.compile(r"^SYNCS")),
    ("MUFU.EX2", re.compile(r"^MUFU\.EX2")),
    ("LDS", re.compile(r"^LDS")),
    ("STS", re.compile(r"^STS")),
    ("ATOMS", re.compile(r"^ATOMS")),
    ("SHFL", re.compile(r"^SHFL")),
    ("BAR", re.compile(r"^BAR")),
    ("LDG", re.compile(r"^LDG")),
    ("STG", re.compile(r"^STG")),
    ("RED/ATOMG", re.compile(r"^(RED|ATOMG)")),
    ("DFMA/DMUL/DADD", re.compile(r"^(DFMA|DMUL|DADD)")),
    ("FFMA", re.compile(r"^FFMA")),
    ("LOP3", re.compile(r"^LOP3")),
    ("POPC", re.compile(r"^POPC")),
]


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return out if len(out) == len(names) else names
    except Exception:
        return names


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)        # argument list
    return name


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "profiles", "r02_sass_histogram.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)")
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = ins.match(line)
        if m:
            cur["__total__"] += 1
            op = m.group(1)
            for col, rx in COLUMNS:
                if rx.match(op):
                    cur[col] += 1
    names = list(kernels)
    pretty = [short(n) for n in demangle(names)]
    rows = sorted(zip(pretty, names), key=lambda t: t[0])
    width = min(72, max(len(p) for p, _ in rows))
    with open(out_path, "w") as f:
        f.write("SASS opcode histogram of scsommerclock_b200/libscs_b200.so (sm_100a), static instruction counts per kernel\n")
        f.write("made by tools/sass_histogram.py from `cuobjdump -sass`; %d kernels\n\n" % len(rows))
        cols = ["total"] + [c for c, _ in COLUMNS]
        f.write("kernel".ljust(width) + "".join(c.rjust(max(8, len(c) + 1)) for c in cols) + "\n")
        tot = collections.Counter()
        for p, n in rows:
            k = kernels[n]
            vals = [k["__total__"]] + [k[c] for c, _ in COLUMNS]
            for c, v in zip(cols, vals):
                tot[c] += v
            f.write(p[:width].ljust(width) + "".join((str(v) if v else ".").rjust(max(8, len(c) + 1)) for c, v in zip(cols, vals)) + "\n")
        f.write("ALL".ljust(width) + "".join(str(tot[c]).rjust(max(8, len(c) + 1)) for c in cols) + "\n")
    print(out_path, len(rows), "kernels")


if __name__ == "__main__":
    main()
