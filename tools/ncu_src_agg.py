#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source function / line.
usage: ncu -i rep --page source --csv --print-source cuda,sass > x.csv; python tools/ncu_src_agg.py x.csv agc_b200/csrc/zstd_enc.cuh"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
srcpath = sys.argv[2]
fp = None
agg = collections.Counter(); samp = collections.Counter()
for r in rows:
    if r and r[0] == "File Path": fp = r[1].split('/')[-1]; continue
    if r and r[0] == "Function Name": continue
    if r and r[0] == "Line No":
        iI = r.index("Instructions Executed"); iS = r.index("# Samples"); continue
    if len(r) > 10 and r[0] != "" and r[2] == "-":
        try: agg[(fp, int(r[0]))] += int(r[iI]); samp[(fp, int(r[0]))] += int(r[iS])
        except ValueError: pass
tot = sum(agg.values()); ts = sum(samp.values())
print("total warp instructions", tot, "samples", ts)
src = open(srcpath).read().split('\n'); base = srcpath.split('/')[-1]
fn_at = {}; cur = None
for i, l in enumerate(src, 1):
    m = re.match(r'^(?:ZE_\w+|template|static|inline)\b[^;]*?(\w+)\s*\([^;]*$', l)
    if m and not l.startswith(' '): cur = m.group(1)
    fn_at[i] = cur
fa = collections.Counter(); fs = collections.Counter()
for (f, l), v in agg.items():
    key = (f, fn_at.get(l) if f == base else None)
    fa[key] += v; fs[key] += samp[(f, l)]
print("--- per function")
for k, v in fa.most_common(30): print("%-40s %12d %5.1f%% inst  %5.1f%% samples" % (k[1] or k[0], v, 100*v/tot, 100*fs[k]/max(ts,1)))
print("--- top lines")
for k, v in agg.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 50):
    print("%s:%d %10d %4.1f%% s%4.1f%% | %s" % (k[0], k[1], v, 100*v/tot, 100*samp[k]/max(ts,1), src[k[1]-1].strip()[:120] if k[0] == base else ''))
