"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel for the LAST
engine solve in the log (launches between two engine_pack_kernel launches); with a second argument, the last solve
that launched a kernel whose name contains it (e.g. 'k1_mask_kernel<4' = the last BATCH solve, not the single-pair
latency probe bench.py ends with)."""
import collections
import csv
import re
import sys


def main(path, must_have=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    allr = []
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = name.replace("void ", "").replace("unnamed>::", "").replace("psulvsb::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        allr.append((name, v))
    idx = [i for i, (n, _) in enumerate(allr) if "engine_pack" in n]
    seg = allr[idx[-2]:idx[-1]] if len(idx) >= 2 else allr
    if must_have and len(idx) >= 2:
        bounds = idx + [len(allr)]
        segs = [allr[bounds[k]:bounds[k + 1]] for k in range(len(bounds) - 1)]
        segs = [g for g in segs if any(must_have in n for n, _ in g)]
        if len(segs) >= 2:
            seg = segs[-2] if len(segs[-1]) < len(segs[-2]) else segs[-1]  # (the capture may cut the last one short)
        elif segs:
            seg = segs[-1]
    tot = sum(v for _, v in seg)
    agg = collections.OrderedDict()
    for n, v in seg:
        a = agg.setdefault(n, [0.0, 0])
        a[0] += v
        a[1] += 1
    print(f"{'kernel':44s} {'launches':>8s} {'total us':>10s} {'share':>7s}")
    for n, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"{n:44s} {c:8d} {v:10.1f} {100 * v / tot:6.1f}%")
    print(f"{'TOTAL (one solve)':44s} {len(seg):8d} {tot:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
