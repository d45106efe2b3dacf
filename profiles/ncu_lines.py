"""Per-source-line hot spots of a captured kernel: joins the ncu SASS page (instructions executed, stall samples)
with nvdisasm -g line info of the in-tree library (same instruction order).
    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep [top]
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr = rows[1]
ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
sass = [(r[1].strip(), float(r[ci] or 0), float(r[si] or 0), float(r[ti] or 0)) for r in rows[2:] if len(r) == len(hdr)]
m = re.search(r"step_kernel<([^>]*)>", kname)
mangled = ("_ZN7shipsim11step_kernelI" + "".join("Li%sE" % v for v in re.findall(r"\(int\)(\d+)", m.group(1))) + "EEvNS_10StepParamsE") if m else None
cubin_prefix = "shipsim_kernels"
mw = re.search(r"window_kernel<\(int\)(\d+), \(int\)(\d+)>", kname)
if mw:
    mangled = "_ZN7shipsim13window_kernelILi%sELi%sEEEvNS_10StepParamsE" % (mw.group(1), mw.group(2))
    cubin_prefix = "shipsim_window"

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("SHIPSIM_LIB") or os.path.join(ROOT, "ship_sim_gym_b200", "libshipsim.so")], cwd=tmp, capture_output=True)
cub = [f for f in glob.glob(os.path.join(tmp, "*.cubin")) if os.path.basename(f).startswith(cubin_prefix)][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
lines = []
inside = False
cur = ("?", 0)
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
print("kernel", kname, "sass rows", len(sass), "disasm instrs", len(lines))
n = min(len(sass), len(lines))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for (txt, ins, smp, tins), loc in zip(sass[:n], lines[:n]):
    a = agg[loc]
    a[0] += ins; a[1] += smp; a[2] += tins
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
src_cache = {}


def src(loc):
    f, l = loc
    if f not in src_cache:
        p = os.path.join(ROOT, "ship_sim_gym_b200", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    s = src_cache[f]
    return s[l - 1].strip()[:100] if 0 < l <= len(s) else ""


print("total warp-instructions %.4g, samples %d" % (tot_i, tot_s))
print("%7s %7s %6s  %s" % ("inst%", "smp%", "lanes", "location"))
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%6.2f%% %6.2f%% %6.1f  %s:%d  %s" % (100 * a[0] / tot_i, 100 * a[1] / tot_s, a[2] / max(a[0], 1), loc[0], loc[1], src(loc)))

if len(sys.argv) > 3:      # full listing in source order, instructions per `norm` (e.g. warp-steps per launch)
    norm = float(sys.argv[3])
    print("\nsource order: warp-instructions per unit (norm %g), stall samples %%, avg active lanes" % norm)
    for loc, a in sorted(agg.items()):
        if a[0] / norm >= 0.5:
            print("%8.1f %6.2f%% %6.1f  %s:%d  %s" % (a[0] / norm, 100 * a[1] / tot_s, a[2] / max(a[0], 1), loc[0], loc[1], src(loc)))

# ---- coarse regions of shipsim_kernels.cu (by marker comments) + device header
import bisect
ksrc = open(os.path.join(ROOT, "ship_sim_gym_b200", "csrc", "shipsim_kernels.cu")).read().splitlines()
marks = [(1, "prologue/misc")]
for i, l in enumerate(ksrc, 1):
    for key, name in (("for (int k = 0; k < p.K", "loop head / action / P frame"), ("handle_discrete_action", "decode"),
                      ("LiDAR.query (models", "lidar: origin + fan box"), ("// pass 1 (cpShape", "lidar pass 1 (planes)"),
                      ("// pass 2, cooperative", "lidar pass 2 (coop rays)"), ("cpSpaceStep: positions", "integrate + sincos"),
                      ("overlap tests at the new pose", "hull box / bank box"), ("cooperative separating-axis", "SAT (coop)"),
                      ("goals: cheap cull", "goals"), ("cpBodyUpdateVelocity", "velocity"), ("determine_reward", "reward/done/stats/reset"),
                      ("---- outputs", "outputs (obs tile, stores)"), ("// episode statistics", "epilogue stats")):
        if key in l:
            marks.append((i, name))
marks.sort()
starts = [m[0] for m in marks]
reg = collections.defaultdict(lambda: [0.0, 0.0])
for loc, a in agg.items():
    if loc[0] == "shipsim_kernels.cu":
        name = marks[bisect.bisect_right(starts, loc[1]) - 1][1]
    else:
        name = loc[0]
    reg[name][0] += a[0]; reg[name][1] += a[1]
print("\nregions (header-file lines are attributed to the header, i.e. inlined helpers / intrinsics):")
for name, a in sorted(reg.items(), key=lambda kv: -kv[1][0]):
    print("%6.2f%% inst %6.2f%% smp   %s" % (100 * a[0] / tot_i, 100 * a[1] / tot_s, name))
