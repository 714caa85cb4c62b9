"""Per-phase view of one kernel of an .ncu-rep: the SASS is cut at the barrier instructions and, for every
piece, instructions executed, stall samples and the stall-reason mix are summed.
usage: python tools/ncu_phases.py rep [kernel-index]"""
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
reasons = [n for n in t["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
print("#", t["name"])
tot = sum(int(r[h["Warp Stall Sampling (All Samples)"]] or 0) for r in t["rows"])
seg = {"first": 0, "n": 0, "ex": 0, "st": 0, "why": {k: 0 for k in reasons}, "ops": {}}
def flush(last, i):
    if seg["n"] == 0: return
    why = sorted(seg["why"].items(), key=lambda kv: -kv[1])[:5]
    ops = sorted(seg["ops"].items(), key=lambda kv: -kv[1])[:6]
    print(f"rows {seg['first']:5d}-{i:5d} instr {seg['ex']:11d} samples {seg['st']:6d} ({100*seg['st']/max(tot,1):5.1f}%)  ends {last[:28]:28s}")
    print("      stalls: " + ", ".join(f"{k[6:]} {100*v/max(seg['st'],1):.0f}%" for k, v in why))
    print("      ops   : " + ", ".join(f"{k} {v/1e6:.1f}M" for k, v in ops))
for i, r in enumerate(t["rows"]):
    src = r[h["Source"]].strip()
    if seg["n"] == 0: seg["first"] = i
    ex = int(r[h["Instructions Executed"]] or 0); st = int(r[h["Warp Stall Sampling (All Samples)"]] or 0)
    seg["n"] += 1; seg["ex"] += ex; seg["st"] += st
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    seg["ops"][op] = seg["ops"].get(op, 0) + ex
    for k in reasons: seg["why"][k] += int(r[h[k]] or 0)
    if "BAR." in src or src.startswith("EXIT") or " EXIT" in src:
        flush(src, i)
        seg = {"first": 0, "n": 0, "ex": 0, "st": 0, "why": {k: 0 for k in reasons}, "ops": {}}
flush("end", len(t["rows"]))
