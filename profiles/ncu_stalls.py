"""Hottest SASS instructions of a captured step kernel by stall samples, with the stall-reason split and source line:
    python profiles/ncu_stalls.py gpurun_out/prof.ncu-rep [top]"""
import csv, glob, io, os, re, subprocess, sys, tempfile, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) == len(hdr)]
m = re.search(r"step_kernel<([^>]*)>", kname)
mangled = ("_ZN7shipsim11step_kernelI" + "".join("Li%sE" % v for v in re.findall(r"\(int\)(\d+)", m.group(1))) + "EEvNS_10StepParamsE") if m else None
cubin_prefix = "shipsim_kernels"
mw = re.search(r"window_kernel<\(int\)(\d+), \(int\)(\d+)>", kname)
if mw:
    mangled = "_ZN7shipsim13window_kernelILi%sELi%sEEEvNS_10StepParamsE" % (mw.group(1), mw.group(2))
    cubin_prefix = "shipsim_window"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("SHIPSIM_LIB") or os.environ.get("SHIPSIM_LIB") or os.path.join(ROOT, "ship_sim_gym_b200", "libshipsim.so")], cwd=tmp, capture_output=True)
cub = [f for f in glob.glob(os.path.join(tmp, "*.cubin")) if os.path.basename(f).startswith(cubin_prefix)][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
lines, inside, cur = [], False, ("?", 0)
for ln in dis:
    if ln.startswith(".text."):
        inside = (ln.strip().rstrip(":") == ".text." + mangled)
        continue
    if not inside:
        continue
    mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if mm:
        cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
tot = collections.Counter()
items = []
for idx, (r, loc) in enumerate(zip(data, lines)):
    smp = float(r[col["# Samples"]] or 0)
    rs = {k: float(r[col[k]] or 0) for k in reasons}
    for k, v in rs.items():
        tot[k] += v
    items.append((smp, idx, loc, r[col["Source"]].strip(), rs))
allsmp = sum(i[0] for i in items) or 1
print("kernel", kname, "total samples", int(allsmp))
print("stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / allsmp) for k, v in tot.most_common(9)))
for smp, idx, loc, txt, rs in sorted(items, reverse=True)[:top]:
    best = sorted(rs.items(), key=lambda kv: -kv[1])[:2]
    print("%5.2f%% #%-5d %-20s %-58s %s" % (100 * smp / allsmp, idx, "%s:%d" % (loc[0][:13], loc[1]), txt[:58],
                                           " ".join("%s=%d" % (k[6:], v) for k, v in best if v)))
