"""Join an `ncu --page source --csv` SASS export with nvdisasm -gi line info: stall samples per CUDA source line.

    python tools/ncu_lines.py <sass_csv> <nvdisasm_gi.sass> <kernel-substring> [source.cu] [min_pct]
Lines are attributed to the OUTERMOST frame that lies in source.cu (the kernel body), so helper functions
inlined into a phase are charged to the phase's call site.
"""
import csv, re, sys
csvf, sassf, kern = sys.argv[1:4]
srcname = sys.argv[4] if len(sys.argv) > 4 else "decode_mega.cu"
minpct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.3
inner = len(sys.argv) > 6 and sys.argv[6] == "inner"   # attribute to the INNERMOST frame in source.cu instead
# --- nvdisasm: instruction offset -> outermost line in srcname
lines = open(sassf).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
off2line = {}
cur = []
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if not m.group(3).strip().startswith("inlined at") and cur and cur[-1][2]:
            cur.append((m.group(1), int(m.group(2)), False))
        else:
            if not cur or not cur[-1][2]:
                cur = []
            cur.append((m.group(1), int(m.group(2)), "inlined at" in m.group(3)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(\S.*);", l)
    if m:
        chain = [c for c in cur if c[0].endswith(srcname)]
        off2line[int(m.group(1), 16)] = (chain[0][1] if inner else chain[-1][1]) if chain else -1
        if cur and not cur[-1][2]:
            pass
rows = list(csv.reader(open(csvf)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[hi + 1][0], 16)
per = {}
tot = toti = 0
for r in rows[hi + 1:]:
    if len(r) <= iI: continue
    off = int(r[0], 16) - base
    ln = off2line.get(off, -2)
    s = int(r[iS] or 0); n = int(r[iI] or 0)
    d = per.setdefault(ln, [0, 0, {}])
    d[0] += s; d[1] += n; tot += s; toti += n
    for i, h in stall_cols:
        v = int(r[i] or 0)
        if v: d[2][h] = d[2].get(h, 0) + v
src = open("/root/repo/music-generation-emotion-adaptive_b200/csrc/" + srcname).read().split("\n")
print("total samples", tot, "instructions", toti)
for ln in sorted(per):
    s, n, st = per[ln]
    if 100 * s / tot >= minpct or 100 * n / toti >= minpct:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        tops = " ".join(f"{k[6:]}:{100*v/max(s,1):.0f}" for k, v in top)
        text = src[ln - 1].strip()[:70] if ln > 0 else "?"
        print(f"{ln:5d} {100*s/tot:5.1f}%smp {100*n/toti:5.1f}%ins  [{tops:40s}] {text}")
