"""Extract dram bytes per launch of the dominant kernel from .ncu-rep files into
profiles/r02_traffic.json (read by bench.py for the `roofline.traffic` fields).
Usage: python profiles/tools/ncu_traffic.py NAME=path.ncu-rep[:kernel-substring] ..."""
import csv, json, os, subprocess, sys

out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "r02_traffic.json")
try:
    res = json.load(open(out_path))
except Exception:
    res = {}
for arg in sys.argv[1:]:
    name, rest = arg.split("=", 1)
    rep, _, pat = rest.partition(":")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    best = None
    for vals in rows[2:]:
        kn = vals[hdr.index("Kernel Name")]
        if pat and pat not in kn:
            continue
        rd = float(vals[hdr.index("dram__bytes_read.sum")].replace(",", ""))
        wr = float(vals[hdr.index("dram__bytes_write.sum")].replace(",", ""))
        units = rows[1]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd *= scale.get(units[hdr.index("dram__bytes_read.sum")], 1)
        wr *= scale.get(units[hdr.index("dram__bytes_write.sum")], 1)
        dur = float(vals[hdr.index("gpu__time_duration.sum")].replace(",", ""))
        if best is None or dur > best[2]:
            best = (kn, rd + wr, dur, rd, wr)
    if best:
        res[name] = int(best[1])
        print(name, best[0][:80], "read", best[3], "write", best[4], "dur", best[2])
json.dump(res, open(out_path, "w"), indent=1, sort_keys=True)
