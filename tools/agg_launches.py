"""Aggregate an ncu launch-list CSV (gpu__time_duration.sum) per kernel for the LAST complete training step."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
idx = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
names = [r[idx["Kernel Name"]] for r in data]
vals = [float(r[idx["Metric Value"]]) for r in data]
unit = data[0][idx["Metric Unit"]]
scale = 1e-3 if unit in ("ns", "nsecond") else 1.0
marks = [i for i, n in enumerate(names) if "loss_fwd_kernel<(int)0>" in n or "image_to_planes" in n and False]
marks = [i for i, n in enumerate(names) if "image_to_planes" in n]
a, b = (marks[-2], marks[-1]) if len(marks) >= 2 else (0, len(names))
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in zip(names[a:b], vals[a:b]):
    k = re.sub(r"\(.*", "", n).replace("void ", "").replace("fsnet::<unnamed>::", "")[:64]
    agg[k][0] += 1
    agg[k][1] += v * scale
tot = sum(v for _, v in agg.values())
print(f"step: {b - a} launches, {tot:.0f} us GPU time (serialised, cold cache)")
print("| us | launches | kernel |\n|---:|---:|---|")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"| {v:.0f} | {c} | {k} |")
