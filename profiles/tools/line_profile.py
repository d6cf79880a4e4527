"""Per-source-line view of an `ncu --set full --import-source on` capture: joins the report's SASS rows (samples,
instructions executed, stall reasons) with the line table nvdisasm prints for the same kernel of the built object
(-lineinfo), instruction by instruction, and prints the lines that hold the most warp samples.
    python profiles/tools/line_profile.py rep.ncu-rep build/k3_rotation.o 'gnc_tls_kernelILi8ELi512' [top=40]
The third argument selects the kernel by a substring of its MANGLED name (as in the object)."""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def line_table(obj, mangled_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, x) for x in os.listdir(tmp) if x.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-c", "-g", cubin], capture_output=True, text=True).stdout
    lines, cur_fn, cur_line, inside = [], None, None, False
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            inside = mangled_sub in m.group(1)
            cur_line = None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            # inlined-at chains: keep the innermost file:line printed last before the instruction
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m:
            lines.append((cur_line, m.group(1).strip()))
    return lines


def main():
    rep, obj, sub = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(obj, sub)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    s = starts[0]
    e = starts[1] if len(starts) > 1 else len(rows)
    hdr, data = rows[s + 1], rows[s + 2:e]
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernel:", rows[s][1][:120])
    print(f"SASS rows {len(data)}, nvdisasm instructions {len(table)}")
    n = min(len(data), len(table))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    S = I = 0.0
    for k in range(n):
        r = data[k]
        key = table[k][0] or ("?", 0)
        a = agg.setdefault(key, [0.0, 0.0, {}])
        sm, ins = f(r[idx["# Samples"]]), f(r[idx["Instructions Executed"]])
        a[0] += sm
        a[1] += ins
        S += sm
        I += ins
        for st in stalls:
            v = f(r[idx[st]])
            if v:
                a[2][st[6:]] = a[2].get(st[6:], 0) + v
    src_cache = {}

    def src(fn, ln):
        if fn not in src_cache:
            p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", fn)
            try:
                src_cache[fn] = open(p).read().splitlines()
            except OSError:
                src_cache[fn] = []
        L = src_cache[fn]
        return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""

    print(f"total samples {S:.0f}, warp instructions {I:.0f}")
    for (fn, ln), (sm, ins, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        tops = ", ".join(f"{k} {100 * v / max(sm, 1):.0f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
        print(f"{fn}:{ln:<5d} samples {100 * sm / S:5.1f}%  instr {100 * ins / I:5.1f}%  [{tops}]  {src(fn, ln)}")


if __name__ == "__main__":
    main()
