"""Instruction mix of a kernel from an .ncu-rep source page: executed warp-instructions per SASS opcode,
and the top SASS lines by stall samples.  usage: python tools/ncu_opmix.py rep [kernel-index]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the csv holds one table per kernel launch, each introduced by a "Kernel Name" row
tables, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; tables.append(cur); continue
    if cur is None or not row: continue
    if cur["hdr"] is None: cur["hdr"] = row; continue
    cur["rows"].append(row)
t = tables[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
h = {n: i for i, n in enumerate(t["hdr"])}
ops, samples = collections.Counter(), collections.Counter()
tot = 0
for r in t["rows"]:
    src = r[h["Source"]].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:2]) if op.startswith(("LDS", "STS", "LDG", "STG")) and "." in op else "")
    n = int(r[h["Instructions Executed"]] or 0)
    ops[op] += n; tot += n
    samples[op] += int(r[h["Warp Stall Sampling (All Samples)"]] or 0)
print("kernel:", t["name"][:100]); print("total warp-instructions:", tot)
st = sum(samples.values())
for op, n in ops.most_common(28):
    print(f"  {op:14s} {n:12d} {100*n/tot:6.2f}%   stall-samples {100*samples[op]/max(st,1):5.1f}%")
