"""Summarise ncu CSV exports (launch list + --set full raw pages) into profiles/<tag>_summary.md."""
import csv, sys, os, json
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
out = []
# ---- launch list ----
rows = list(csv.reader(open(os.path.join(src, f"launches_{tag}.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]; data = rows[hi + 1:]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = {}
for r in data:
    if len(r) <= vi: continue
    d = per.setdefault(int(r[ii]), {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
agg = {}
tot_t = 0.0
for k in sorted(per):
    d = per[k]
    nm = d["name"].split("(")[0].split("::")[-1].replace("void ", "").strip()
    a = agg.setdefault(nm, {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
    a["n"] += 1; a["ns"] += d.get("gpu__time_duration.sum", 0.0)
    a["rd"] += d.get("dram__bytes_read.sum", 0.0); a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot_t += d.get("gpu__time_duration.sum", 0.0)
out.append(f"## Launch list of one steady-state step ({len(per)} launches of our kernels, {tot_t/1e3:.1f} us summed; ncu serialises and runs cold-cache: compare SHARES)\n")
out.append("| kernel | launches | time us | share | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|")
for nm, a in sorted(agg.items(), key=lambda x: -x[1]["ns"]):
    out.append(f"| {nm} | {a['n']} | {a['ns']/1e3:.1f} | {a['ns']/tot_t*100:.1f}% | {a['rd']/1e6:.1f} | {a['wr']/1e6:.1f} |")
traffic = {nm: {"launches": a["n"], "dram_bytes": a["rd"] + a["wr"], "time_ns": a["ns"]} for nm, a in agg.items()}
json.dump(traffic, open(os.path.join("profiles", f"traffic_{tag}.json"), "w"), indent=1)
# ---- full captures ----
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM rd"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp insts")]
for part in ("gemm", "conv3", "post", "dec", "dw"):
    f = os.path.join(src, f"prof_{part}_{tag}_raw.csv")
    if not os.path.exists(f): continue
    rows = list(csv.reader(open(f)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out.append(f"\n## ncu --set full: {part}\n")
    out.append("| kernel | " + " | ".join(w[1] for w in want) + " |\n|---|" + "---|" * len(want))
    for r in data:
        nm = r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1].replace("void ", "")
        cells = []
        for key, _ in want:
            if key in hdr:
                i = hdr.index(key); cells.append(f"{r[i]} {units[i]}".strip())
            else: cells.append("-")
        out.append(f"| {nm} | " + " | ".join(cells) + " |")
open(os.path.join("profiles", f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
