"""Condense an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv <command>`) into
per-kernel totals and shares.  usage: python tools/ncu_launch_summary.py gpurun_out/x.csv "<command>" > profiles/rNN_x_summary.csv"""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if r]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; iK = H.index("Kernel Name"); iV = H.index("Metric Value"); iU = H.index("Metric Unit"); iM = H.index("Metric Name")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= iV or r[iM] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "").strip()
    v = float(r[iV].replace(",", ""))
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iU], 1e-6)
    tot[name] += ms; cnt[name] += 1
T = sum(tot.values())
print(f"# ncu launch list of `{sys.argv[2] if len(sys.argv) > 2 else '?'}` (per-launch times are cold-cache and serialised: compare SHARES, not absolutes)")
print(f"# total kernel time {T:.1f} ms over {sum(cnt.values())} launches")
print("kernel,launches,total_ms,share")
for k, v in tot.most_common():
    print(f"{k},{cnt[k]},{v:.3f},{v / T:.4f}")
