"""Rank source lines of one kernel by executed warp instructions.

    python scripts/ncu_hot_lines.py REPORT.ncu-rep CUBIN MANGLED_KERNEL [top]

Joins ncu's SASS page (per-instruction counts) with nvdisasm's line info (build with -lineinfo).
Inlined code is attributed to the innermost source line."""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep, cubin, fun = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
line_of = {}
cur = None
inside = False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        inside = ln.strip() == f".text.{fun}:"
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ie, te, sa = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = None
agg = defaultdict(lambda: [0, 0, 0])
tot = tott = 0
for r in rows[hi + 1:]:
    try:
        addr = int(r[0], 16); v = int(r[ie]); t = int(r[te]); s = int(r[sa])
    except (ValueError, IndexError):
        continue
    if base is None:
        base = addr
    key = line_of.get(addr - base, (None, ""))[0]
    a = agg[key]
    a[0] += v; a[1] += t; a[2] += s
    tot += v; tott += t
print(f"total warp instr {tot:,}  thread instr {tott:,}  avg active lanes {tott / max(tot, 1):.1f}")
srcs = {}
def src(key):
    if not key: return ""
    f, l = key
    if f not in srcs:
        import glob
        c = glob.glob(f"/root/repo/**/{f}", recursive=True)
        srcs[f] = open(c[0]).read().splitlines() if c else []
    L = srcs[f]
    return L[l - 1].strip()[:100] if 0 < l <= len(L) else ""
for key, (v, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * v / tot:5.2f}%  lanes {t / max(v, 1):4.1f}  samples {s:6d}  {key}  {src(key)}")
