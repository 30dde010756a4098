"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list (last training step only)."""
import collections, csv, re, sys
f = open(sys.argv[1]).read().splitlines()
i = [k for k, l in enumerate(f) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(f[i:]))
names = [r['Kernel Name'] for r in rows]
vals = [float(r['Metric Value'].replace(',', '')) for r in rows]
# last step = from the last launch of the first forward kernel (first conv) to the end
starts = [k for k, n in enumerate(names) if 'fc1_cov_kernel' in n or 'first_conv_kernel<0' in n]
half = starts[-1] if starts else len(rows) // 2
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in zip(names[half:], vals[half:]):
    n = re.sub(r'\(.*', '', n).replace('void ', '').replace('ub::', '')
    agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print("kernel,launches,total_us,share")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:70]},{c},{t/1e3:.1f},{t/tot:.4f}")
print(f"TOTAL,{len(rows)-half},{tot/1e3:.1f},1.0")
