"""Sum the DRAM bytes of the forward launches in an `ncu --metrics ... --csv` log (tools/profile.sh) and record them in
profiles/forward_traffic.json, which bench.py reads for `roofline.traffic`:
    python tools/forward_traffic.py gpurun_out/fwd_metrics_<tag>.csv profiles/<tag>_forward_per_launch.md [batch size]"""
import csv
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
src, source_md = sys.argv[1], sys.argv[2]
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 32
size = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
rows = [r for r in csv.reader(l for l in open(src).read().splitlines() if l.startswith('"'))]
hdr = rows[0]
ii = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = 0.0
ids = set()
for r in rows[1:]:
    if r[ii["Metric Name"]] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[ii["Metric Value"]].replace(",", "")) * scale[r[ii["Metric Unit"]]]
        ids.add(r[ii["ID"]])
out = REPO / "profiles" / "forward_traffic.json"
recs = json.loads(out.read_text()) if out.exists() else []
recs = [x for x in recs if (x["batch"], x["size"], tuple(x["dilations"])) != (batch, size, (1, 2, 4, 8, 16))]
recs.append({"batch": batch, "size": size, "dilations": [1, 2, 4, 8, 16], "dram_bytes": tot, "launches": len(ids),
             "source": source_md})
out.write_text(json.dumps(recs, indent=1) + "\n")
print(f"{len(ids)} launches, {tot / 1e9:.2f} GB -> {out}")
