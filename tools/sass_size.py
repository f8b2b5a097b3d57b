"""Static SASS size of one kernel broken down by the source function its instructions come from (nvdisasm line info).
usage: sass_size.py lib.so mangled_kernel_name      (no GPU needed; the instruction cache budget is 32 KB)"""
import collections, os, re, subprocess, sys, tempfile
lib, fn = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
csrc = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc")
funcs = {}   # file -> sorted [(line, name)]
for f in os.listdir(csrc):
    rows = []
    for i, ln in enumerate(open(os.path.join(csrc, f), errors="replace").read().splitlines(), 1):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:__device__|__global__|__host__)[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", ln)
        if m and not ln.strip().startswith("//"):
            rows.append((i, m.group(1)))
    funcs[f] = rows
def owner(f, line):
    best = "?"
    for l, name in funcs.get(f, []):
        if l <= line: best = name
        else: break
    return best
secs = [fn] + [m for m in re.findall(r"\.text\.(\S+):", txt) if m != fn and ("cold" in m or "ddiv_ieee" in m)]
for sname in secs:
    if f".text.{sname}:" not in txt: continue
    sec = txt.split(f".text.{sname}:")[1]
    nxt = sec.find("//--------------------- .text.")
    sec = sec[:nxt] if nxt > 0 else sec
    cur, cnt, total = ("?", 0), collections.Counter(), 0
    for ln in sec.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln):
            cnt[(cur[0], owner(cur[0], cur[1]))] += 1; total += 1
    print(f"== {sname}: {total} instructions = {total * 16 / 1024:.1f} KB")
    for (f, name), v in sorted(cnt.items(), key=lambda kv: -kv[1])[:25]:
        print(f"   {v:5d}  {v * 16 / 1024:5.1f} KB  {f}:{name}")
