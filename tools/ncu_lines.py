"""Join ncu's SASS-level samples with nvdisasm line info: samples / instructions per source line.
usage: ncu_lines.py report.ncu-rep lib.so mangled_kernel_name [top_n]"""
import csv, io, re, subprocess, sys, tempfile, os, collections
rep, lib, fn = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
sec = txt.split(f".text.{fn}:")[1]
nxt = sec.find("//--------------------- .text.")
sec = sec[:nxt] if nxt > 0 else sec
addr2line, cur = {}, ("?", 0)
for ln in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", ln)
    if m: addr2line[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
recs = [r for r in rows[2:] if len(r) >= len(hdr)]
base = int(recs[0][col["Address"]], 16)
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for r in recs:
    a = int(r[col["Address"]], 16) - base
    key = addr2line.get(a, ("?", 0))
    agg[key][0] += float(r[col["# Samples"]] or 0)
    agg[key][1] += float(r[col["Instructions Executed"]] or 0)
    agg[key][2] += float(r[col["Thread Instructions Executed"]] or 0)
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
src = {}
for f in set(k[0] for k in agg):
    for d in ("petershirleyraytracer_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)
        if os.path.exists(p): src[f] = open(p).read().splitlines()
print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s} {'thr/inst':>8s}  source")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src.get(f, [""] * (l + 1))[l - 1].strip()[:90] if f in src and 0 < l <= len(src[f]) else ""
    print(f"{f + ':' + str(l):28s} {100 * v[0] / ts:6.2f} {100 * v[1] / ti:6.2f} {v[2] / max(v[1], 1):8.1f}  {text}")
