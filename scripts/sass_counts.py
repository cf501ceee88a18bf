"""SASS mnemonic counts per kernel of libalpharat_cuda.so (B200_PROFILING.md: the mnemonics that prove tcgen05 / TMA).
    python scripts/sass_counts.py > profiles/r2_sass_counts.md"""
import re, subprocess, sys
from collections import OrderedDict
so = "alpharat_b200/libalpharat_cuda.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDG", "STG", "LDS", "STS", "LDL", "STL", "REDUX", "SHFL", "ATOM", "MUFU", "FFMA", "IMAD", "HMMA"]
cur = None
tab = OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        tab[cur] = {k: 0 for k in MN}
        tab[cur]["_total"] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1).split(".")[0]
        tab[cur]["_total"] += 1
        if op in tab[cur]:
            tab[cur][op] += 1
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
print("# SASS mnemonic counts per kernel (`cuobjdump -sass alpharat_b200/libalpharat_cuda.so`, sm_100a)\n")
print("UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk copy\nengine, mbarrier complete_tx), UTMALDG = cp.async.bulk.tensor (tensor-map TMA, not used: the operands are pre-swizzled\non the host and moved as flat bulk copies), SYNCS = mbarrier ops.\n")
cols = ["_total"] + MN
print("| kernel | " + " | ".join(c.strip("_") for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for k, v in tab.items():
    print("| `" + demangle(k) + "` | " + " | ".join(str(v[c]) for c in cols) + " |")
