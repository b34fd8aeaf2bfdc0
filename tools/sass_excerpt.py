"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md), from the
built library:  python tools/sass_excerpt.py > profiles/r2_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flowdec_b200", "libflowdec_b200.so")
PAT = re.compile(r"\b(UTCHMMA[A-Z0-9_.]*|UTCMMA[A-Z0-9_.]*|UTCBAR[A-Z0-9_.]*|LDTM[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|UTMASTG[A-Z0-9_.]*|"
                 r"UTMAPF[A-Z0-9_.]*|SYNCS[A-Z0-9_.]*|FFMA2|FADD2|MUFU\.TANH|MUFU\.COS|ELECT|NANOSLEEP[A-Z0-9_.]*)")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, per = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn)
            per[fn] = collections.Counter()
            continue
        if fn:
            for t in PAT.findall(line):
                per[fn][t.rstrip(".")] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): mnemonic counts per kernel")
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA load/store,")
    print("# UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, FFMA2 = packed fp32x2 FMA")
    for fn, c in per.items():
        if any(k.startswith(("UTC", "LDTM", "UTMA")) for k in c):
            print(f"{fn}: " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))


if __name__ == "__main__":
    sys.exit(main())
