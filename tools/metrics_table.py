"""ncu --csv --metrics log -> one row per launch: python tools/metrics_table.py <csv> [out.md]"""
import csv, sys, re, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]).read().splitlines() if l.startswith('"'))]
hdr = rows[0]; ii = {h: i for i, h in enumerate(hdr)}
launches = collections.OrderedDict()
for r in rows[1:]:
    key = r[ii["ID"]]
    name = re.sub(r"void |\(unnamed\)::|unnamed>::|dc::|<unnamed>::|\(.*\)$", "", r[ii["Kernel Name"]]).strip()
    d = launches.setdefault(key, {"kernel": name})
    v = r[ii["Metric Value"]].replace(",", "")
    d[r[ii["Metric Name"]]] = (float(v), r[ii["Metric Unit"]])
def g(d, k, scale=1.0, fmt="{:.1f}"):
    if k not in d: return "-"
    v, u = d[k]
    if u in ("ns", "nsecond"): v /= 1e6
    elif u in ("us", "usecond"): v /= 1e3
    elif u == "Gbyte": v *= 1e3
    elif u == "Kbyte": v /= 1e3
    elif u == "byte": v /= 1e6
    return fmt.format(v * scale)
cols = [("ms", "gpu__time_duration.sum", "{:.3f}"), ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "{:.1f}"),
        ("tc-smem%", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "{:.1f}"),
        ("lsu%", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "{:.1f}"),
        ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "{:.1f}"), ("L2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "{:.1f}"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "{:.1f}"),
        ("dramR MB", "dram__bytes_read.sum", "{:.0f}"), ("dramW MB", "dram__bytes_write.sum", "{:.0f}"),
        ("L2->SM MB", "l1tex__m_xbar2l1tex_read_bytes.sum", "{:.0f}"), ("regs", "launch__registers_per_thread", "{:.0f}"),
        ("smem MB", "launch__shared_mem_per_block_dynamic", "{:.3f}"), ("grid", "launch__grid_size", "{:.0f}")]
names = sys.argv[3].split(",") if len(sys.argv) > 3 else None
out = ["| # | launch | kernel | " + " | ".join(c[0] for c in cols) + " |", "|" + "---|" * (len(cols) + 3)]
for n, (k, d) in enumerate(launches.items()):
    lab = names[n] if names and n < len(names) else ""
    out.append(f"| {n} | {lab} | `{d['kernel']}` | " + " | ".join(g(d, c[1], 1.0, c[2]) for c in cols) + " |")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2 and sys.argv[2] != "-": open(sys.argv[2], "w").write(text + "\n")
