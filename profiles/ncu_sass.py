"""SASS of given source lines of shipsim_kernels.cu with per-instruction execution counts from an .ncu-rep:
    python profiles/ncu_sass.py gpurun_out/prof.ncu-rep <norm> <line> [<line> ...]
(the in-tree libshipsim.so must be the build the capture was taken from)."""
import csv, glob, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, norm = sys.argv[1], float(sys.argv[2])
want = set(int(x) for x in sys.argv[3:])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr = rows[1]
ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
sass = [(r[1].strip(), float(r[ci] or 0), float(r[si] or 0), float(r[ti] or 0)) for r in rows[2:] if len(r) == len(hdr)]
m = re.search(r"step_kernel<([^>]*)>", kname)
mangled = ("_ZN7shipsim11step_kernelI" + "".join("Li%sE" % v for v in re.findall(r"\(int\)(\d+)", m.group(1))) + "EEvNS_10StepParamsE") if m else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("SHIPSIM_LIB") or os.path.join(ROOT, "ship_sim_gym_b200", "libshipsim.so")], cwd=tmp, capture_output=True)
cub = [f for f in glob.glob(os.path.join(tmp, "*.cubin")) if os.path.basename(f).startswith("shipsim_kernels")][0]
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
for idx, ((txt, ins, smp, tins), loc) in enumerate(zip(sass, lines)):
    if (loc[0] == "shipsim_kernels.cu" and loc[1] in want) or not want:
        print("%5d %-22s %8.2f %5.1f %6d  %s" % (idx, "%s:%d" % (loc[0][:14], loc[1]), ins / norm, tins / max(ins, 1), smp, txt[:120]))
