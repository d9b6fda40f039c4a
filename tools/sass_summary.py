#!/usr/bin/env python
"""profiles/sass_summary.txt: per kernel of libgrakel_b200.so, the counts of the SASS mnemonics that prove the
Blackwell paths (tcgen05 MMA / TMEM loads / TMA loads and stores / cluster barriers), from `cuobjdump -sass`.
Runs in the build container (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "grakel_b200", "libgrakel_b200.so")
PAT = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS",
       "SYNCS", "ATOMG", "ATOM", "RED", "ATOMS", "BAR", "CCTL", "LDGSTS", "MUFU", "DFMA", "DADD", "DMUL", "HMMA", "IMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            base = op.split(".")[0]
            kernels[cur][base] += 1
            if op.startswith("UTCHMMA.2CTA") or ".2CTA" in op and base == "UTCHMMA":
                kernels[cur]["UTCHMMA.2CTA"] += 1
    demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    lines = ["# SASS mnemonic counts per kernel of grakel_b200/libgrakel_b200.so (cuobjdump -sass, sm_100a)",
             "# UTCHMMA = tcgen05.mma kind::f16/tf32, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA load/store, UTCBAR = tcgen05.commit,",
             "# SYNCS = mbarrier ops, ATOMG/RED = global atomics", ""]
    for (name, cnt), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm)
        keys = [k for k in PAT if cnt.get(k)]
        lines.append("%-70s instr %6d  %s" % (short[:70], cnt["_total"], "  ".join("%s=%d" % (k, cnt[k]) for k in keys)))
    path = os.path.join(ROOT, "profiles", "sass_summary.txt")
    open(path, "w").write("\n".join(lines) + "\n")
    print("\n".join(l for l in lines if "UTC" in l or "UTMA" in l))


if __name__ == "__main__":
    main()
