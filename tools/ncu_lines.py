"""Join an ncu `--page source --csv` dump (per-SASS-instruction metrics) with nvdisasm -g line
info and aggregate executed instructions / stall samples per CUDA source line.

usage: python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top-N]
"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, so, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
lines_by_idx = []
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True,
                         text=True).stdout
    insec = False
    cur = ("?", 0, "?", 0)
    for ln in txt.splitlines():
        if ln.startswith("//----") and ".text." in ln:
            insec = kname in ln
            continue
        if not insec:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inner = (os.path.basename(m.group(1)), int(m.group(2)))
            chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            outer = (os.path.basename(chain[-1][0]), int(chain[-1][1])) if chain else inner
            cur = inner + outer
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines_by_idx.append(cur)
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                        text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
print("sass instrs: ncu %d, nvdisasm %d" % (len(data), len(lines_by_idx)))
agg_in = collections.defaultdict(lambda: [0, 0, 0])
agg_out = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for i, r in enumerate(data):
    if i >= len(lines_by_idx):
        break
    ex = int(float(r[ci["Instructions Executed"]] or 0))
    sm = int(float(r[ci["# Samples"]] or 0))
    lsb = int(float(r[ci["stall_long_sb"]] or 0))
    inner = lines_by_idx[i][:2]; outer = lines_by_idx[i][2:]
    for a, k in ((agg_in, inner), (agg_out, outer)):
        a[k][0] += ex; a[k][1] += sm; a[k][2] += lsb
    tot[0] += ex; tot[1] += sm; tot[2] += lsb
print("total warp-instr %d samples %d long_sb %d" % tuple(tot))
for title, a in (("INNERMOST line", agg_in), ("OUTERMOST (kernel-level) line", agg_out)):
    print("==== by %s: file:line  instr%%  samples%%  long_sb%%" % title)
    for k, v in sorted(a.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-22s %6.2f %6.2f %6.2f" % ("%s:%d" % k, 100 * v[0] / tot[0], 100 * v[1] / tot[1],
                                          100 * v[2] / max(1, tot[2])))

# ---- grouped by model function (line ranges of hx_model.cuh / hx_kernels.cu) ----
import bisect
def groups_for(path):
    starts = []
    for n, ln in enumerate(open(path), 1):
        m = re.match(r"(?:template.*\n)?(?:__device__|__global__|static).*?(\w+)\(", ln)
        if m and ("__device__" in ln or "__global__" in ln):
            starts.append((n, m.group(1)))
    return starts
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gm = groups_for(os.path.join(root, "hector_b200/csrc/hx_model.cuh"))
g = collections.defaultdict(lambda: [0, 0, 0])
for (f, line), v in agg_in.items():
    if f == "hx_model.cuh":
        i = bisect.bisect_right([s for s, _ in gm], line) - 1
        name = gm[i][1] if i >= 0 else "?"
    else:
        name = f + ":kernel-body"
    for k in range(3):
        g[name][k] += v[k]
print("==== by function: instr% samples% long_sb%")
for k, v in sorted(g.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %6.2f %6.2f %6.2f" % (k, 100 * v[0] / tot[0], 100 * v[1] / tot[1], 100 * v[2] / max(1, tot[2])))

# ---- static SASS size by function ----
st = collections.Counter()
for (f, line, _, _) in lines_by_idx:
    if f == "hx_model.cuh":
        i = bisect.bisect_right([s for s, _ in gm], line) - 1
        st[gm[i][1] if i >= 0 else "?"] += 1
    else:
        st[f] += 1
print("==== static SASS instruction count by function (x16 B)")
for k, v in st.most_common(20):
    print("%-28s %6d" % (k, v))

# ---- cold but bulky: source lines with many SASS instructions that almost never execute ----
# (candidates to move out of the hot instruction stream: the year body is I-cache bound)
per_line = collections.defaultdict(lambda: [0, 0])  # static count, executed
for i, r in enumerate(data):
    if i >= len(lines_by_idx):
        break
    k = lines_by_idx[i][:2]
    per_line[k][0] += 1
    per_line[k][1] += int(float(r[ci["Instructions Executed"]] or 0))
hot = max(v[1] / max(1, v[0]) for v in per_line.values())
print("==== cold code inside the kernel: file:line  static instrs  executions per instr / hottest")
cold = [(k, v) for k, v in per_line.items() if v[0] >= 12 and v[1] / v[0] < 0.02 * hot]
for k, v in sorted(cold, key=lambda kv: -kv[1][0])[:25]:
    print("%-22s %5d  %.4f" % ("%s:%d" % k, v[0], v[1] / v[0] / hot))
print("cold static instrs total: %d of %d" % (sum(v[0] for _, v in cold), sum(v[0] for v in per_line.values())))
