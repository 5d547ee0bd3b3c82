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

# ---- instruction mix of the whole library and the lines that prove the bulk copy / mbarrier /
# shuffle / compare-and-swap instructions of the chosen kernel (profiles/r02_sass_static_*.txt)
sass = subprocess.run(["cuobjdump", "-sass", os.path.abspath(so)], capture_output=True, text=True).stdout
mix = collections.Counter()
proof = collections.Counter()
inside = False
for ln in sass.splitlines():
    if "Function :" in ln:
        inside = kname in ln
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if not m:
        continue
    mix[m.group(1)] += 1
    if inside and re.search(r"UBLKCP|SYNCS|SHFL|ATOMG|ATOM\.|MEMBAR|BAR\.|CCTL\.E\.PF", ln):
        proof[re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip()] += 1
print("\n==== instruction mix of the whole cubin (cuobjdump -sass %s | mnemonic counts)" % so)
for k, v in mix.most_common(40):
    print("%7d %s" % (v, k))
print("\n==== lines proving TMA bulk copy / mbarrier / shuffles / CAS / L2 prefetch in %s" % kname)
for k, v in proof.most_common(40):
    print("%7d %s" % (v, k))
