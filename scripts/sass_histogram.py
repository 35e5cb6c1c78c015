"""Opcode histogram of every kernel in yael_b200/libyael_b200.so (cuobjdump -sass), the evidence
that the shipped cubins are hand-written sm_100a code: tcgen05 MMAs (UTCHMMA f16/tf32, UTCQMMA
f8f6f4), TMA loads and stores (UTMALDG, UTMASTG, UBLKCP), cp.async (LDGSTS), TMEM loads (LDTM), mbarrier waits (SYNCS), packed FP32 (FFMA2),
three-input min/max (FMNMX3), population counts (POPC).

    python scripts/sass_histogram.py > profiles/r2_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "yael_b200", "libyael_b200.so")
KEY = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "LDGSTS", "UBLKCP", "UTCBAR", "LDTM", "STTM", "SYNCS",
       "ELECT", "POPC", "FFMA2", "FMNMX3", "FMNMX", "HFMA2", "HMNMX2", "LOP3", "ATOMS", "ATOMG", "RED",
       "LDG", "STG", "LDS", "STS", "SHFL", "MATCH", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    funcs = collections.OrderedDict()
    cur = None
    for ln in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = funcs.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["__total__"] += 1
    dm = demangle(list(funcs))
    print("# cuobjdump -sass yael_b200/libyael_b200.so -- architectures:", ", ".join(arch))
    tot = collections.Counter()
    for c in funcs.values():
        tot.update(c)
    print("# whole library: %d kernels, %d SASS instructions" % (len(funcs), tot["__total__"]))
    print("# " + "  ".join("%s=%d" % (k, sum(v for o, v in tot.items() if o.startswith(k))) for k in KEY))
    print()
    for name, c in funcs.items():
        keys = [(k, sum(v for o, v in c.items() if o == k or o.startswith(k + "."))) for k in KEY]
        keys = [(k, v) for k, v in keys if v]
        short = re.sub(r"\(.*", "", dm.get(name, name))
        print("%-72s %6d instr  %s" % (short[:72], c["__total__"], "  ".join("%s=%d" % kv for kv in keys)))


if __name__ == "__main__":
    sys.exit(main())
