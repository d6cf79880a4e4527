"""Where a kernel's warp time goes, by straight-line SASS region: reads `ncu -i rep --page source --csv
--print-source sass` (a `--set full --import-source on` capture) and prints, per kernel, the stall-reason totals and
every region (instructions between two branches / barriers) holding more than THR of the warp samples, with its
share of the executed instructions and its top stall reasons.  A region with many samples and few instructions is
latency nobody hides (how the K1 slow path and the GNC parking swap were found in round 1).
Usage: stall_regions.py rep.ncu-rep [thr=0.02] [kernel-substring]"""
import csv
import io
import re
import subprocess
import sys


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main():
    rep = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for n, s in enumerate(starts):
        e = starts[n + 1] if n + 1 < len(starts) else len(rows)
        name = rows[s][1]
        if want not in name:
            continue
        hdr, data = rows[s + 1], rows[s + 2:e]
        idx = {h: i for i, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = {st: sum(f(r[idx[st]]) for r in data) for st in stalls}
        T = sum(tot.values()) or 1.0
        print("==", name[:140])
        print("   samples %d: " % T + ", ".join(f"{k[6:]} {100 * v / T:.1f}%" for k, v in
                                                  sorted(tot.items(), key=lambda x: -x[1]) if v / T > 0.01))
        regions, cur = [], []
        for r in data:
            cur.append(r)
            if re.search(r"\bBRA\b|BAR\.|EXIT|CALL|RET", r[idx["Source"]]):
                regions.append(cur)
                cur = []
        if cur:
            regions.append(cur)
        S = sum(f(r[idx["# Samples"]]) for r in data) or 1.0
        I = sum(f(r[idx["Instructions Executed"]]) for r in data) or 1.0
        for reg in regions:
            sm = sum(f(r[idx["# Samples"]]) for r in reg)
            ins = sum(f(r[idx["Instructions Executed"]]) for r in reg)
            if sm / S > thr:
                det = {st[6:]: int(sum(f(r[idx[st]]) for r in reg)) for st in stalls}
                det = dict(sorted(det.items(), key=lambda x: -x[1])[:4])
                worst = max(reg, key=lambda r: f(r[idx["# Samples"]]))
                print(f"   {reg[0][idx['Address']][-5:]}-{reg[-1][idx['Address']][-5:]} {len(reg):4d} instr  samples {100 * sm / S:5.1f}%"
                      f"  executed {100 * ins / I:5.1f}%  {det}  worst: {worst[idx['Source']].strip()[:60]}")


if __name__ == "__main__":
    main()
