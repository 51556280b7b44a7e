"""Aggregate ncu warp-stall samples per CUDA source line (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
def num(x):
    try: return int(x)
    except ValueError: return 0
hdr = None; out = []; cur = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; ix = {}
        for i, h in enumerate(hdr): ix.setdefault(h, i)
        keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or not r or not r[0].isdigit(): continue
    st = {k[6:]: num(r[ix[k]]) for k in keys}
    out.append((num(r[ix["# Samples"]]), cur, int(r[0]), r[1].strip(), num(r[ix["Instructions Executed"]]), st))
tot = sum(o[0] for o in out); print("total samples", tot)
for n, f, ln, src, inst, st in sorted(out, reverse=True)[:top]:
    t3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{f:20s}:{ln:4d} {n:6d} {100*n/max(tot,1):5.1f}% inst={inst:9d} {src[:60]:60s} {t3}")
