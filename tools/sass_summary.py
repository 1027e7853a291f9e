"""Per-kernel SASS opcode summary of libsegofa_b200.so (cuobjdump -sass): the instructions that prove the tcgen05 / TMEM /
TMA path -- UTCHMMA (tcgen05.mma, .2CTA = cta_group::2), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UTMAREDG (TMA tensor
load / store / reduce), UTCBAR (tcgen05.commit), UBLKCP (bulk copy), SYNCS (mbarrier)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ifseg_b200", "libsegofa_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "UBLKCP", "SYNCS", "MUFU", "HMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["_total"] += 1
        base = op.split(".")[0]
        if base in OPS:
            cur[base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            cur["UTCHMMA.2CTA"] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    fam = collections.OrderedDict()
    for mangled, name in zip(per, names):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("sgf::", "")
        key = re.sub(r"<.*", "", short)
        f = fam.setdefault(key, dict(n=0, c=collections.Counter()))
        f["n"] += 1
        f["c"].update(per[mangled])
    print("# SASS opcode counts per kernel family of ifseg_b200/libsegofa_b200.so (sm_100a), summed over the template instantiations")
    print("# (tools/sass_summary.py; n = instantiations)")
    print(f"{'kernel':40s} {'n':>3s} {'instrs':>8s} " + " ".join(f"{o:>12s}" for o in OPS))
    tot = collections.Counter()
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["c"]["UTCHMMA"]):
        print(f"{k[:40]:40s} {f['n']:3d} {f['c']['_total']:8d} " + " ".join(f"{f['c'][o]:12d}" for o in OPS))
        tot.update(f["c"])
    print(f"{'TOTAL':40s} {sum(f['n'] for f in fam.values()):3d} {tot['_total']:8d} " + " ".join(f"{tot[o]:12d}" for o in OPS))


if __name__ == "__main__":
    main()
