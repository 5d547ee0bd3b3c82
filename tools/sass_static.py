"""static SASS instruction count per source function / line of one kernel (nvdisasm -g line info)
usage: python tools/sass_static.py <lib.so> <kernel-substring> [top]"""
import os, re, subprocess, sys, tempfile, collections, bisect
so, kname = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = []
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    insec = False
    cur = ("?", 0)
    for ln in txt.splitlines():
        if ln.startswith("//----") and ".text." in ln:
            insec = kname in ln
            continue
        if not insec:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
print("kernel", kname, "static SASS instructions:", len(lines), "=", len(lines) * 16 // 1024, "KB")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
starts = []
for n, ln in enumerate(open(os.path.join(root, "hector_b200/csrc/hx_model.cuh")), 1):
    m = re.match(r"(?:__device__|__global__|static).*?(\w+)\(", ln)
    if m and "__device__" in ln:
        starts.append((n, m.group(1)))
st = collections.Counter()
for f, line in lines:
    if f == "hx_model.cuh":
        i = bisect.bisect_right([s for s, _ in starts], line) - 1
        st[starts[i][1] if i >= 0 else "?"] += 1
    else:
        st[f] += 1
for k, v in st.most_common(top):
    print("%-28s %6d" % (k, v))
