"""Dump the SASS of kernel #idx of an .ncu-rep with executed counts and stall samples. usage: ncu_sass.py rep [idx] > file"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
tables, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; tables.append(cur); continue
    if cur is None or not row: continue
    if cur["hdr"] is None: cur["hdr"] = row; continue
    cur["rows"].append(row)
t = tables[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
h = {n: i for i, n in enumerate(t["hdr"])}
print("#", t["name"])
for i, r in enumerate(t["rows"]):
    print(f"{i:5d} {int(r[h['Instructions Executed']] or 0):10d} {int(r[h['Warp Stall Sampling (All Samples)']] or 0):6d}  {r[h['Source']].strip()}")
