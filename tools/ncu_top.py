"""Top stall sites of an ncu capture: python tools/ncu_top.py <file.ncu-rep> [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
start = rows.index(hdr)
si, ci, ai = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Address")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0     # which captured kernel
blocks, cur = [], None
for r in rows:
    if r == hdr:
        cur = []; blocks.append(cur); continue
    if cur is not None and len(r) == len(hdr): cur.append(r)
data = blocks[which]
tot = sum(int(r[ci] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
idx = {id(r): i for i, r in enumerate(data)}
for r in sorted(data, key=lambda r: -int(r[ci] or 0))[:n]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{int(r[ci] or 0):8d} {100*int(r[ci] or 0)/max(tot,1):5.1f}%  #{idx[id(r)]:5d} {r[si][:70]:70s} {st}")
