#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv --log-file F` launch list (this library's kernels only).
usage: summarize_ncu_launches.py launches.csv "header text" > profiles/launches_rN_summary.txt"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    name = r[ik]
    if "at::" in name or not any(s in name for s in ("mw::", "logmel", "<unnamed>")):     # torch's own kernels (fills, copies) out
        continue
    us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("mw::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    tot[short][0] += us
    tot[short][1] += 1
total = sum(v[0] for v in tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print(f"# {sum(v[1] for v in tot.values())} launches of this library, {total / 1e3:.3f} ms summed (cold-cache, serialised: compare SHARES, not absolutes)")
for k, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{us / 1e3:9.3f} ms {100 * us / total:5.1f}%  n={n:5d} avg {us / n:9.2f} us  {k}")
