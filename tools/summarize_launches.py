import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ik, im, iv, iu, ig = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("Grid Size")
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[hdr + 1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    name = r[ik].split("(")[0] + " grid" + r[ig]
    agg[name][r[im]].append(v * scale)
tot = sum(sum(m["gpu__time_duration.sum"]) for m in agg.values())
print(f"{'kernel':70s} {'n':>4s} {'avg us':>9s} {'share':>6s} {'GB r+w':>8s} {'GB/s':>8s}")
for k, m in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
    t = m["gpu__time_duration.sum"]
    b = [a + c for a, c in zip(m.get("dram__bytes_read.sum", [0] * len(t)), m.get("dram__bytes_write.sum", [0] * len(t)))]
    avg_t, avg_b = sum(t) / len(t), sum(b) / max(len(b), 1)
    print(f"{k:70s} {len(t):4d} {avg_t:9.1f} {100 * sum(t) / tot:5.1f}% {avg_b / 1e9:8.3f} {avg_b / avg_t / 1e3:8.0f}")
