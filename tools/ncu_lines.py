"""Attribute the warp-stall samples of an `ncu --page source --csv` export (SASS rows) to CUDA source lines.
  nvdisasm -g -c <cubin> > dis.txt ;  ncu -i rep.ncu-rep --page source --csv > src.csv
  python tools/ncu_lines.py dis.txt src.csv <mangled-substring> [top]
nvdisasm annotates every instruction with `//## File "...", line N` (innermost inlined frame); rows are joined by the
instruction offset inside the function."""
import csv
import collections
import re
import sys

dis, src, key = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
line_of = {}
infn = False
cur = None
for ln in open(dis, errors="replace"):
    if ln.startswith(".text."):
        infn = key in ln
        cur = None
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src, newline="")))
hdr = next(r for r in rows if r and r[0] == "Address")
ia, isamp, iexec = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
base = int(body[0][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r in body:
    off = int(r[ia], 16) - base
    where, _ = line_of.get(off, (None, ""))
    s = int(r[isamp] or 0)
    tot += s
    a = agg[where]
    a[0] += s
    a[1] += int(r[iexec] or 0)
    for i, h in stall_cols:
        v = int(r[i] or 0)
        if v:
            a[2][h[6:]] += v
print(f"total samples {tot}, SASS rows {len(body)}, mapped {sum(1 for r in body if (int(r[ia],16)-base) in line_of)}")
for where, (s, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = " ".join(f"{k}={v}" for k, v in st.most_common(3))
    print(f"{100.0*s/tot:6.2f}%  {s:7d}  exec={ex:10d}  {where}  {tops}")
byfile = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for where, (s, ex, st) in agg.items():
    f = where[0] if where else None
    byfile[f][0] += s
    byfile[f][1] += ex
    byfile[f][2].update(st)
print("-- by file")
for f, (s, ex, st) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"{100.0*s/tot:6.2f}%  exec={ex:10d}  {f}  " + " ".join(f"{k}={v}" for k, v in st.most_common(5)))
allst = collections.Counter()
for _, (_, _, st) in agg.items():
    allst.update(st)
print("-- stalls: " + " ".join(f"{k}={100.0*v/tot:.1f}%" for k, v in allst.most_common(8)))
print(f"-- executed warp instructions: {sum(a[1] for a in agg.values())}")
