"""Turn the ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarize_profiles.py <tag>          # reads gpurun_out/launches_<tag>.csv, prof_*_<tag>.ncu-rep
"""
import csv, json, subprocess, sys, collections, re
from pathlib import Path

tag = sys.argv[1]
G = Path("gpurun_out"); P = Path("profiles"); P.mkdir(exist_ok=True)

def short(name):
    name = re.sub(r"void |\(unnamed\)::|unnamed>::|dc::|<unnamed>::|\(.*\)$", "", name)
    return name.strip()

# ---- launch list: per-kernel totals and shares over the captured window
lf = G / f"launches_{tag}.csv"
if lf.exists():
    rows = [r for r in csv.reader(l for l in lf.read_text().splitlines() if l.startswith('"'))]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    tot = collections.OrderedDict(); n = collections.Counter()
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")); u = r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v
        k = short(r[ki]); tot[k] = tot.get(k, 0.0) + us; n[k] += 1
    total = sum(tot.values())
    out = [f"# ncu launch list, tag {tag}: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` under",
           "# ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache and serialised:",
           "# compare SHARES, not absolutes).  launches captured: %d, summed device time %.1f ms" % (len(rows) - 1, total / 1e3),
           "kernel,launches,total_ms,share"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        out.append(f"{k},{n[k]},{v/1e3:.3f},{v/total:.4f}")
    (P / f"{tag}_launch_shares.csv").write_text("\n".join(out) + "\n")
    print("\n".join(out[:16]))

# ---- full captures: key metrics per captured launch
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
lines = [f"# ncu --set full --clock-control none captures, tag {tag} (one row block per captured launch)"]
for rep in sorted(G.glob(f"prof_*_{tag}.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3: continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        lines.append(f"\n## {rep.name}: {short(r[hdr.index('Kernel Name')])}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w); lines.append(f"{w} = {r[i]} {units[i]}")
(P / f"{tag}_ncu_full.txt").write_text("\n".join(lines) + "\n")
print("\n".join(lines[:60]))
