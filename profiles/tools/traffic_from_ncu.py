"""dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels in `ncu --set full` reports -> profiles/ncu_traffic.json
(read by bench.py for `roofline.traffic`, which cannot be measured in-process).
    python profiles/tools/traffic_from_ncu.py <registrations per launch> rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    per_launch = int(sys.argv[1])
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ncu_traffic.json")
    table = {}
    if os.path.exists(out_path):
        table = json.load(open(out_path))
    for rep in sys.argv[2:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        for r in rows[2:]:
            name = re.sub(r"[<(].*", "", r[ik]).split("::")[-1].strip()
            b = float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
            table[name] = {"bytes_per_registration": b / per_launch,
                           "source": f"ncu capture profiles/{os.path.basename(rep).replace('.ncu-rep', '').replace('r2_', 'r2_ncu_')}.txt "
                                     f"(B = {per_launch})"}
    json.dump(table, open(out_path, "w"), indent=1)
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    main()
