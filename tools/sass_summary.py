#!/usr/bin/env python
"""Per-kernel SASS instruction summary of libconformer_b200.so (cuobjdump -sass; runs without a GPU).

Counts the mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduction store), and the legacy HMMA (mma.sync),
plus FFMA / FFMA2 and MUFU for the CUDA-core kernels.  Writes a markdown table to stdout."""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "conformer_b200", "lib", "libconformer_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCMMA_OTHER", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "FFMA", "FFMA2", "MUFU", "SHFL", "ATOM/RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        c = per[cur]
        c["total"] += 1
        if base == "UTCHMMA":
            c["UTCHMMA"] += 1
            if ".2CTA" in op:
                c["UTCHMMA.2CTA"] += 1                          # tcgen05.mma.cta_group::2 (issued for a CTA pair)
        elif base.startswith("UTC") and "MMA" in base:
            c["UTCMMA_OTHER"] += 1
        elif base in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "FFMA", "FFMA2", "MUFU", "SHFL"):
            c[base] += 1
        elif base in ("ATOM", "ATOMG", "RED", "ATOMS"):
            c["ATOM/RED"] += 1
    names = list(per)
    if names:
        dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, dm)) if len(dm) == len(names) else {}
    print("| kernel | SASS instrs | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    tot = collections.Counter()
    for n, c in sorted(per.items(), key=lambda kv: demangle.get(kv[0], kv[0])):
        name = demangle.get(n, n)
        name = re.sub(r"\(.*", "", name).replace("cfm::(anonymous namespace)::", "").replace("cfm::", "")
        name = re.sub(r"^void ", "", name)
        print(f"| `{name[:70]}` | {c['total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
        tot.update(c)
    print(f"| **total ({len(per)} kernels)** | {tot['total']} | " + " | ".join(str(tot[k]) for k in KEYS) + " |")


if __name__ == "__main__":
    sys.exit(main())
