"""Aggregate an .ncu-rep's per-instruction stall samples by CUDA source line.
python tools/ncu_lines.py report.ncu-rep lib.so mangled-kernel-substring [top]
(SASS offsets -> lines from `nvdisasm -g` of the cubin inside the .so; needs -lineinfo)"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, so, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of, cur, inside = {}, None, False
for l in dis.splitlines():
    if l.startswith("//---") and ".text." in l:
        inside = kern in l
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name' and data:
        break
    if len(r) == len(hdr) and r[0] != 'Address':
        data.append(r)
base = int(data[0][ci['Address']], 16)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
reasons = [h for h in hdr if h.startswith('stall_')]
for r in data:
    off = int(r[ci['Address']], 16) - base
    key = line_of.get(off)
    n = int(r[ci['# Samples']]); tot += n
    a = agg[key]; a[0] += n; a[1] += int(r[ci['Instructions Executed']])
    for h in reasons:
        a[2][h] += int(r[ci[h]])
print(f"{tot} samples, {len(data)} SASS instructions, {len(agg)} source lines")
srcs = {}
for key, (n, ex, rs) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if key:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "riskaversetrajopt_b200", "csrc", key[0])
        if os.path.exists(path):
            srcs.setdefault(path, open(path).read().splitlines())
            text = srcs[path][key[1] - 1].strip()[:70]
    top3 = ", ".join(f"{k[6:]}={v}" for k, v in rs.most_common(3))
    print(f"{100*n/tot:5.1f}%  {n:6d}  inst {ex:9d}  {key}  [{top3}]  {text}")
