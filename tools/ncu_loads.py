#!/usr/bin/env python
"""tools/ncu_loads.py <report> <kernel substr>: executed global-load instructions per warp, by source line (serialised memory round trips)."""
import csv, re, subprocess, sys, tempfile, os, collections
rep, kern = sys.argv[1], sys.argv[2]
nwarps = float(sys.argv[3]) if len(sys.argv) > 3 else 2048.0
so = "dcmrta_b200/libdcmrta_b200.so"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
off2line = {}; cur = None
for l in dis[start + 1:]:
    if l.startswith("//---") or (l.startswith(".text.") and l.rstrip().endswith(":")):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
kidx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kern.split("ILi")[0].lstrip("_Z0123456789") in r[1]]
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address" and (not kidx or i > kidx[0]))
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ("Address", "Source", "Instructions Executed", "# Samples", "Thread Instructions Executed", "stall_long_sb")}
base = None
agg = collections.defaultdict(collections.Counter)
for r in rows[hdr_i + 1:]:
    if not r or not r[0].startswith("0x"):
        break
    a = int(r[0], 16); base = a if base is None else base
    src = r[ci["Source"]]
    key = off2line.get(a - base, ("?", 0))
    n = int(float(r[ci["Instructions Executed"]] or 0))
    agg[key]["inst"] += n
    agg[key]["lsb"] += int(float(r[ci["stall_long_sb"]] or 0))
    if re.search(r"\bLDG|\bLD\.E|\bLDL", src):
        agg[key]["ld"] += n; agg[key]["ldthreads"] += int(float(r[ci["Thread Instructions Executed"]] or 0))
    if re.search(r"\bSTG|\bRED|\bATOM", src):
        agg[key]["st"] += n
tl = sum(c["ld"] for c in agg.values()); ts = sum(c["st"] for c in agg.values()); ti = sum(c["inst"] for c in agg.values())
print(f"per warp: {ti/nwarps:.0f} instructions, {tl/nwarps:.1f} load instr, {ts/nwarps:.1f} store/red instr")
print("%-18s %5s %9s %6s %8s" % ("file", "line", "ld/warp", "lanes", "long_sb"))
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["ld"])[:40]:
    if c["ld"]:
        print("%-18s %5d %9.2f %6.1f %8d" % (key[0], key[1], c["ld"] / nwarps, c["ldthreads"] / max(1, c["ld"]), c["lsb"]))
