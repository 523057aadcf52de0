"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libotk.so (tcgen05 MMA = UTCHMMA / UTCQMMA..., TMA loads /
stores = UTMALDG / UTMASTG, tensor-memory loads / stores = LDTM / STTM, tcgen05 barriers = UTCBAR) plus the MUFU / DFMA /
FFMA mix of the SIMT kernels.     python profiles/tools/sass_summary.py > profiles/sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "ot-vae-lightning_b200", "csrc", "libotk.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "MUFU.EX2", "DFMA", "FFMA", "HFMA2",
        "LDG.E.128", "ATOM", "RED"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        c = counts[cur]
        c["total"] += 1
        for k in KEYS:
            if k == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    c[k] += 1
            elif op.startswith(k):
                c[k] += 1
    names = demangle(list(counts))
    print("# SASS summary of libotk.so (sm_100a)\n")
    print("`cuobjdump -sass ot-vae-lightning_b200/csrc/libotk.so`, instruction counts per kernel (static, not executed counts).")
    print("UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / st")
    print("(tensor memory), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops.\n")
    cols = KEYS + ["total"]
    print("| kernel | " + " | ".join(cols) + " |")
    print("|---|" + "---|" * len(cols))
    tot = collections.Counter()
    for fn, c in sorted(counts.items(), key=lambda kv: -(kv[1]["UTCHMMA"] * 1000 + kv[1]["total"])):
        short = re.sub(r"\(.*", "", names.get(fn, fn)).replace("void ", "").replace("otk::", "")
        print(f"| `{short[:70]}` | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
        tot.update(c)
    print("| **all kernels** | " + " | ".join(str(tot[k]) for k in cols) + " |")
    print(f"\n{len(counts)} kernels.")


if __name__ == "__main__":
    main()
