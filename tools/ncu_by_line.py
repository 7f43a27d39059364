#!/usr/bin/env python
"""tools/ncu_by_line.py -- attribute an ncu SASS source page to CUDA source lines via nvdisasm line info.
usage: ncu_by_line.py <report.ncu-rep> <kernel mangled substring> [lib.so]"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, kern = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else "dcmrta_b200/libdcmrta_b200.so"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# find function section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
off2line = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith("//---") or (l.startswith(".text.") and l.rstrip().endswith(":")):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        # innermost location first; keep the full inline chain
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
# first kernel only
kidx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kern.split("ILi")[0].lstrip("_Z0123456789") in r[1]]
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address" and (not kidx or i > kidx[0]))
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ("Address", "Source", "Instructions Executed", "# Samples", "Thread Instructions Executed", "stall_no_inst", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving")}
base = None
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[hdr_i + 1:]:
    if not r or not r[0].startswith("0x"):
        break
    a = int(r[0], 16)
    base = a if base is None else base
    loc = off2line.get(a - base, ("?", 0, ""))
    key = (loc[0], loc[1])
    for n in ("Instructions Executed", "# Samples", "Thread Instructions Executed", "stall_no_inst", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving"):
        v = int(float(r[ci[n]] or 0))
        agg[key][n] += v
        tot[n] += v
print("total inst", tot["Instructions Executed"], "samples", tot["# Samples"], "thread-inst/inst %.1f" % (tot["Thread Instructions Executed"] / max(1, tot["Instructions Executed"])))
print("%-22s %6s %8s %6s %6s | no_inst long_sb short_sb wait branch" % ("file", "line", "inst%", "smp%", "lanes"))
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:int(os.environ.get("TOP", "45"))]:
    print("%-22s %6d %7.2f%% %5.1f%% %6.1f | %5d %5d %5d %5d %5d" % (key[0], key[1], 100 * c["Instructions Executed"] / tot["Instructions Executed"], 100 * c["# Samples"] / tot["# Samples"],
          c["Thread Instructions Executed"] / max(1, c["Instructions Executed"]), c["stall_no_inst"], c["stall_long_sb"], c["stall_short_sb"], c["stall_wait"], c["stall_branch_resolving"]))

# ---- per-function buckets (source function containing the innermost line) ----
def func_table(path):
    tab = []
    try:
        for i, l in enumerate(open(path), 1):
            m = re.search(r"__(?:device|global)__.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
            if m and not l.strip().startswith("//"):
                tab.append((i, m.group(1)))
    except OSError:
        pass
    return tab
tabs = {f: func_table(os.path.join("dcmrta_b200/csrc", f)) for f in ("dcm_thread.cuh", "dcm_kernels.cu")}
fagg = collections.defaultdict(collections.Counter)
for (f, ln), c in agg.items():
    name = f
    for (s, n) in tabs.get(f, []):
        if s <= ln:
            name = n
    fagg[name].update(c)
print("\nper function (instructions per env-step assume 65536 envs):")
for name, c in sorted(fagg.items(), key=lambda kv: -kv[1]["Instructions Executed"]):
    print("%-28s inst %6.1f%% (%6.0f/step)  samples %5.1f%%  lanes %5.1f" % (name, 100 * c["Instructions Executed"] / tot["Instructions Executed"], c["Instructions Executed"] / 65536,
          100 * c["# Samples"] / tot["# Samples"], c["Thread Instructions Executed"] / max(1, c["Instructions Executed"])))
